// fluxnet.hpp — the reference's network-rendering protocol over the GPU worker.
//
//   fluxcore/src/workers.rs:106-110   enum NetworkWorkerRequest { SetJob(Box<Job>), WorkUnit(WorkUnit), Done }
//   fluxcore/src/workers.rs:118-245   NetworkWorker: the manager's end (connect, read WorkerInfo, SetJob, two units
//                                     in flight, one RenderEvent back per unit, Done)
//   flux-node/src/main.rs:21-94       handle_client: the node's end (write WorkerInfo, then serve requests)
//   flux-node/src/main.rs:96-111      run_server: one client at a time
//   fluxcore/src/manager.rs:16-28     RenderEvent::RowsReady(WorkUnitResult{work_unit, rows: Vec<Vec<Color>>})
//   fluxcore/src/job.rs:10,40-62      JobID(usize, usize), WorkUnit, JobConfiguration, Job{id, scene_data, config}
//
// Messages are CBOR items written back to back on one TCP connection (serde_cbor 0.9 to_writer /
// StreamDeserializer; see cbor.hpp for the forms).  `NodeServer` is what `fluxb200-node` runs in place of
// flux-node: an unmodified `flux -n host:port` manager can drive the GPUs of this box.  `NetworkWorker` is the
// manager's end, used by `fluxb200 -n host:port` and by the tests.
//
// The Rust reference cannot be built in this image, so the wire forms follow serde_cbor's documented behaviour
// and are pinned only by hand-derived byte vectors (tests/test_net_protocol.py), not by the reference itself.
// The one point on which serde_cbor releases differ — the form of enum variants — is handled without a guess:
// the node answers in the form the manager's requests arrive in, and the manager's end defaults to the 0.9 form.
#pragma once
#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "fluxhost.hpp"

namespace flux {

struct JobID { uint64_t allocator_id = 0, id = 0; };                       // job.rs:10
struct Job { JobID id; SceneData scene_data; JobConfiguration config; };   // job.rs:58-62

namespace net {

constexpr const char *DEFAULT_PORT = "2000";   // constants.rs:6

// How a newtype enum variant travels: serde_cbor < 0.10 (the reference pins 0.9.0) writes the 2-array
// [name, content]; 0.10 and later write the one-entry map {name: content} and read both.
enum class EnumForm { Array, Map };

struct Request {   // NetworkWorkerRequest, workers.rs:106-110
    enum Kind { SetJob, Unit, Done } kind = Done;
    EnumForm form = EnumForm::Array;   // the form this request arrived in (a unit variant has none: Array)
    Job job;
    WorkUnit unit{0, 0, 0, 0};
};

// ---- wire forms (pure functions; tests call them through `fluxb200-node --encode/--decode`) ----
std::string encode_worker_info(uint64_t num_threads);                        // manager.rs:221-224
std::string encode_set_job(const Job &job, EnumForm form = EnumForm::Array);
std::string encode_work_unit(const WorkUnit &unit, EnumForm form = EnumForm::Array);
std::string encode_done();
std::string encode_rows_ready(const WorkUnitResult &r, uint32_t width, EnumForm form = EnumForm::Array);   // RenderEvent::RowsReady
// one item from the front of `data`; `used` = its length.  Throw flux::Error on malformed input.
Request decode_request(const void *data, size_t n, size_t *used = nullptr);
uint64_t decode_worker_info(const void *data, size_t n, size_t *used = nullptr);
WorkUnitResult decode_rows_ready(const void *data, size_t n, uint32_t *width, size_t *used = nullptr);

// ---- the node's end ----
class NodeServer {
  public:
    // binds and listens at once (port "0" = any free port, see port()); throws flux::Error
    NodeServer(GpuWorker &worker, const std::string &host, const std::string &port);
    ~NodeServer();
    uint16_t port() const { return port_; }
    // run_server (flux-node/src/main.rs:96-111): accept and serve clients one at a time; with max_clients > 0
    // return after that many.  A client that breaks the protocol is dropped with a message, like handle_client's Err.
    void serve(uint64_t max_clients = 0);
    void stop();   // from another thread: makes serve() return after the current client

  private:
    void handle_client(int fd, const std::string &peer);
    GpuWorker &worker_;
    int listen_fd_ = -1;
    uint16_t port_ = 0;
    std::atomic<bool> stop_{false};
};

// The work units of one job, handed out one at a time to whoever asks: the shared bounded(1) queue the
// reference's workers pull from (manager.rs:100,156-162).
class UnitQueue {
  public:
    explicit UnitQueue(const Job &job);   // work_units() of the job, in row order
    bool pop(WorkUnit &u);                // thread-safe; false when every unit has been handed out

  private:
    std::mutex mu_;
    std::vector<WorkUnit> units_;
    size_t next_ = 0;
};

// ---- the manager's end (workers.rs:118-245) ----
class NetworkWorker {
  public:
    explicit NetworkWorker(const std::string &endpoint, EnumForm form = EnumForm::Array);   // "host" or "host:port"
    ~NetworkWorker();
    NetworkWorker(const NetworkWorker &) = delete;
    NetworkWorker &operator=(const NetworkWorker &) = delete;
    WorkerInfo info() const { return info_; }
    // one job over the connection: SetJob, two units in flight, one RowsReady per unit, Done (which ends the
    // connection: flux-node returns from handle_client on Done)
    Image render_job(const Job &job);
    // the same, taking units from a queue shared with other workers and writing rows into a shared image
    void run(const Job &job, UnitQueue &queue, Image &img);

  private:
    int fd_ = -1;
    WorkerInfo info_;
    std::string endpoint_;
    EnumForm form_;
};

// One job over several nodes (`flux -n a -n b`): every node gets the job, all pull units from one queue.  With
// `local`, the GPUs of this box pull from the same queue (the reference's LocalWorker beside its NetworkWorkers,
// flux/src/main.rs:43-60; `-L` leaves it out).  Which worker renders a unit is a matter of timing, as in the
// reference; with the same seed on every worker the frame does not depend on it.
Image render_job_on_nodes(const std::vector<std::string> &endpoints, const Job &job, EnumForm form = EnumForm::Array,
                          std::vector<WorkerInfo> *infos = nullptr, GpuWorker *local = nullptr);

}  // namespace net
}  // namespace flux
