// host_slices.h — an index range in slices on several host threads (flux_set_scene's per-triangle passes, the BVH
// builder).  Every caller's slices are independent and write disjoint outputs, so the result does not depend on the
// number of threads — or on whether a thread could be had at all: std::thread's constructor throws std::system_error when
// the process is out of threads, and then the slice runs inline (no exception may leave a joinable thread behind or
// cross the C ABI).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <memory>
#include <system_error>
#include <thread>
#include <utility>
#include <vector>

// A vector whose resize() leaves trivially constructible elements unwritten: the hundred-megabyte arrays of a
// million-triangle scene are then first touched (and their pages faulted in) by the slices that fill them, on several
// threads, instead of being zeroed by one.  Every element is written before it is read — the arrays that go to the
// device are hashed whole by flux_bvh_hash, so a byte left behind would show as a tree that differs from run to run.
template <class T> struct flux_noinit_alloc : std::allocator<T> {
    template <class U> struct rebind { using other = flux_noinit_alloc<U>; };
    flux_noinit_alloc() = default;
    template <class U> flux_noinit_alloc(const flux_noinit_alloc<U> &) {}
    template <class U> void construct(U *p) { ::new (static_cast<void *>(p)) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
template <class T> using flux_raw_vector = std::vector<T, flux_noinit_alloc<T>>;

template <class F> inline void flux_spawn_or_run(std::vector<std::thread> &pool, F f) {
    try {
        pool.emplace_back(f);
    } catch (const std::system_error &) {
        f();
    }
}

inline void flux_in_slices(uint32_t count, const std::function<void(uint32_t, uint32_t)> &body) {
    const uint32_t nthreads = count >= (1u << 16) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
    if (nthreads == 1) {
        body(0, count);
        return;
    }
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nthreads; t++) {
        const uint32_t lo = (uint32_t)((uint64_t)count * t / nthreads), hi = (uint32_t)((uint64_t)count * (t + 1) / nthreads);
        flux_spawn_or_run(th, [&body, lo, hi] { body(lo, hi); });
    }
    for (auto &x : th) x.join();
}
