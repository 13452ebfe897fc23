// fluxnet.cpp — see fluxnet.hpp.
#include "fluxnet.hpp"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

#include "cbor.hpp"
#include "node.hpp"

namespace flux {
namespace net {

using cbor::Writer;
using detail::Node;

// ------------------------------------------------------------------------------------------------
// Serialize: the field names and order of the reference's derive(Serialize) types
// ------------------------------------------------------------------------------------------------
namespace {

void put_vec3(Writer &w, const Vec3 &v) {   // nalgebra Vector3 / Point3: the 3 coordinates as a sequence
    w.array(3);
    for (double c : v) w.f64(c);
}

void put_color(Writer &w, const Vec3 &c) {   // color.rs:12-16
    w.map(3);
    w.key("r").f64(c[0]);
    w.key("g").f64(c[1]);
    w.key("b").f64(c[2]);
}

void put_material(Writer &w, const MaterialData &m) {   // shapes.rs:42-82, externally tagged
    if (auto *x = std::get_if<MatteData>(&m)) {
        w.variant("Matte");
        w.map(3);
        w.key("diffuse_color");
        put_color(w, x->diffuse_color);
        w.key("ambient_color");
        put_color(w, x->ambient_color);
        w.key("diffuse_coefficient").f64(x->diffuse_coefficient);
    } else if (auto *e = std::get_if<EmissiveData>(&m)) {
        w.variant("Emissive");
        w.map(2);
        w.key("color");
        put_color(w, e->color);
        w.key("power").f64(e->power);
    } else if (auto *r = std::get_if<ReflectiveData>(&m)) {
        w.variant("Reflective");
        w.map(2);
        w.key("reflect_amount").f64(r->reflect_amount);
        w.key("reflect_color");
        put_color(w, r->reflect_color);
    } else {
        const auto &g = std::get<GlossyReflectiveData>(m);
        w.variant("GlossyReflective");
        w.map(3);
        w.key("reflect_amount").f64(g.reflect_amount);
        w.key("reflect_color");
        put_color(w, g.reflect_color);
        w.key("reflect_exponent").f64(g.reflect_exponent);
    }
}

void put_shape(Writer &w, const ShapeData &s) {   // scene.rs:71-74, shapes.rs:18-37 (+ extensions)
    if (auto *sp = std::get_if<SphereData>(&s)) {
        w.variant("Sphere");
        w.map(4);
        w.key("center");
        put_vec3(w, sp->center);
        w.key("radius").f64(sp->radius);
        w.key("material");
        put_material(w, sp->material);
        w.key("invert").boolean(sp->invert);
    } else if (auto *pl = std::get_if<PlaneData>(&s)) {
        w.variant("Plane");
        w.map(3);
        w.key("point");
        put_vec3(w, pl->point);
        w.key("normal");
        put_vec3(w, pl->normal);
        w.key("material");
        put_material(w, pl->material);
    } else if (auto *t = std::get_if<TriangleData>(&s)) {
        w.variant("Triangle");
        w.map(4);
        w.key("v0");
        put_vec3(w, t->v0);
        w.key("v1");
        put_vec3(w, t->v1);
        w.key("v2");
        put_vec3(w, t->v2);
        w.key("material");
        put_material(w, t->material);
    } else if (auto *m = std::get_if<MeshData>(&s)) {
        w.variant("Mesh");
        w.map(3);
        w.key("vertices").array(m->vertices.size());
        for (const Vec3 &v : m->vertices) put_vec3(w, v);
        w.key("faces").array(m->faces.size());
        for (const auto &f : m->faces) {
            w.array(3);
            for (int64_t i : f) w.uint((uint64_t)i);
        }
        w.key("material");
        put_material(w, m->material);
    } else if (auto *r = std::get_if<RectangleData>(&s)) {
        w.variant("Rectangle");
        w.map(4);
        w.key("corner");
        put_vec3(w, r->corner);
        w.key("edge_a");
        put_vec3(w, r->edge_a);
        w.key("edge_b");
        put_vec3(w, r->edge_b);
        w.key("material");
        put_material(w, r->material);
    } else {
        const auto &b = std::get<BoxData>(s);
        w.variant("Box");
        w.map(3);
        w.key("min");
        put_vec3(w, b.min);
        w.key("max");
        put_vec3(w, b.max);
        w.key("material");
        put_material(w, b.material);
    }
}

void put_scene(Writer &w, const SceneData &sd) {   // scene.rs:42-49
    w.map(6);
    w.key("scene_name").text(sd.scene_name);
    w.key("output_settings").map(3);
    w.key("image_width").uint(sd.output_settings.image_width);
    w.key("image_height").uint(sd.output_settings.image_height);
    w.key("pixel_size").f64(sd.output_settings.pixel_size);
    w.key("background");
    put_color(w, sd.background);
    w.key("shapes").array(sd.shapes.size());
    for (const ShapeData &s : sd.shapes) put_shape(w, s);
    w.key("camera_settings").map(3);
    w.key("eye");
    put_vec3(w, sd.camera_settings.eye);
    w.key("look_at");
    put_vec3(w, sd.camera_settings.look_at);
    w.key("up");
    put_vec3(w, sd.camera_settings.up);
    w.key("camera_data").map(4);
    w.key("zoom_factor").f64(sd.camera_data.zoom_factor);
    w.key("view_plane_distance").f64(sd.camera_data.view_plane_distance);
    w.key("focal_distance").f64(sd.camera_data.focal_distance);
    w.key("lens_radius").f64(sd.camera_data.lens_radius);
}

void put_job_id(Writer &w, uint64_t allocator_id, uint64_t id) {   // tuple struct JobID(usize, usize)
    w.array(2);
    w.uint(allocator_id);
    w.uint(id);
}

void put_work_unit(Writer &w, const WorkUnit &u) {   // job.rs:40-44
    w.map(3);
    w.key("row_start").uint(u.row_start);
    w.key("row_end").uint(u.row_end);
    w.key("job_id");
    put_job_id(w, u.job_allocator_id, u.job_id);
}

// ------------------------------------------------------------------------------------------------
// Deserialize
// ------------------------------------------------------------------------------------------------
const Node &field(const Node &m, const char *key, const char *what) {
    const Node *n = m.find(key);
    if (!n) throw Error(std::string(what) + ": missing field `" + key + "`");
    return *n;
}

uint64_t as_u64(const Node &n, const char *what) {
    if (n.kind == Node::Scalar && n.bin == Node::U64) return n.u;
    throw Error(std::string(what) + ": invalid type: expected an unsigned integer");
}

uint32_t as_u32(const Node &n, const char *what) {
    const uint64_t v = as_u64(n, what);
    if (v > 0xFFFFFFFFull) throw Error(std::string(what) + ": value does not fit 32 bits");
    return (uint32_t)v;
}

double as_double(const Node &n, const char *what) {
    if (n.kind == Node::Scalar && n.bin == Node::F64) return n.f;
    if (n.kind == Node::Scalar && n.bin == Node::U64) return (double)n.u;
    if (n.kind == Node::Scalar && n.bin == Node::I64) return (double)(int64_t)n.u;
    throw Error(std::string(what) + ": invalid type: expected f64");
}

// a struct arrives as a map; serde's derived visitor also takes the fields in order as a sequence
const Node &struct_field(const Node &s, size_t index, const char *key, const char *what) {
    if (s.kind == Node::Seq) {
        if (index >= s.seq.size()) throw Error(std::string(what) + ": invalid length " + std::to_string(s.seq.size()));
        return s.seq[index];
    }
    if (s.kind != Node::Map) throw Error(std::string(what) + ": expected a map");
    return field(s, key, what);
}

void job_id_from(const Node &n, uint64_t &allocator_id, uint64_t &id) {
    if (n.kind != Node::Seq || n.seq.size() != 2) throw Error("JobID: expected a sequence of 2 integers");
    allocator_id = as_u64(n.seq[0], "JobID.0");
    id = as_u64(n.seq[1], "JobID.1");
}

WorkUnit work_unit_from(const Node &n) {
    WorkUnit u{0, 0, 0, 0};
    u.row_start = as_u32(struct_field(n, 0, "row_start", "WorkUnit"), "WorkUnit.row_start");
    u.row_end = as_u32(struct_field(n, 1, "row_end", "WorkUnit"), "WorkUnit.row_end");
    job_id_from(struct_field(n, 2, "job_id", "WorkUnit"), u.job_allocator_id, u.job_id);
    return u;
}

Job job_from(const Node &n) {
    Job j;
    job_id_from(struct_field(n, 0, "id", "Job"), j.id.allocator_id, j.id.id);
    j.scene_data = detail::scene_from_node(struct_field(n, 1, "scene_data", "Job"));
    const Node &c = struct_field(n, 2, "config", "Job");
    j.config.sample_root = as_u32(struct_field(c, 0, "sample_root", "JobConfiguration"), "sample_root");
    j.config.max_trace_depth = as_u32(struct_field(c, 1, "max_trace_depth", "JobConfiguration"), "max_trace_depth");
    j.config.rows_per_work_unit = as_u32(struct_field(c, 2, "rows_per_work_unit", "JobConfiguration"), "rows_per_work_unit");
    return j;
}

// an enum value: "Variant" (unit), {"Variant": content} (serde_cbor >= 0.9), or the older ["Variant", content]
void variant_of(const Node &n, const char *what, std::string &tag, const Node *&content, EnumForm *form = nullptr) {
    content = nullptr;
    if (form) *form = n.kind == Node::Map ? EnumForm::Map : EnumForm::Array;
    if (n.kind == Node::Scalar && n.bin == Node::Text) {
        tag = n.scalar;
    } else if (n.kind == Node::Map && n.map.size() == 1) {
        tag = n.map[0].first;
        content = &n.map[0].second;
    } else if (n.kind == Node::Seq && !n.seq.empty() && n.seq.size() <= 2 && n.seq[0].kind == Node::Scalar && n.seq[0].bin == Node::Text) {
        tag = n.seq[0].scalar;
        if (n.seq.size() == 2) content = &n.seq[1];
    } else {
        throw Error(std::string(what) + ": expected an enum variant");
    }
}

Request request_from(const Node &n) {
    std::string tag;
    const Node *content;
    Request r;
    variant_of(n, "NetworkWorkerRequest", tag, content, &r.form);
    if (tag == "Done") {
        r.kind = Request::Done;
    } else if (tag == "WorkUnit") {
        if (!content) throw Error("NetworkWorkerRequest::WorkUnit: missing content");
        r.kind = Request::Unit;
        r.unit = work_unit_from(*content);
    } else if (tag == "SetJob") {
        if (!content) throw Error("NetworkWorkerRequest::SetJob: missing content");
        r.kind = Request::SetJob;
        r.job = job_from(*content);
    } else {
        throw Error("unknown variant `" + tag + "`, expected one of `SetJob`, `WorkUnit`, `Done`");
    }
    return r;
}

uint64_t worker_info_from(const Node &n) { return as_u64(struct_field(n, 0, "num_threads", "WorkerInfo"), "WorkerInfo.num_threads"); }

WorkUnitResult rows_ready_from(const Node &n, uint32_t *width) {
    std::string tag;
    const Node *content;
    variant_of(n, "RenderEvent", tag, content);
    if (tag != "RowsReady" || !content) throw Error("RenderEvent: expected RowsReady, got `" + tag + "`");
    WorkUnitResult r{work_unit_from(struct_field(*content, 0, "work_unit", "WorkUnitResult")), {}};
    const Node &rows = struct_field(*content, 1, "rows", "WorkUnitResult");
    if (rows.kind != Node::Seq) throw Error("WorkUnitResult.rows: expected a sequence");
    uint32_t w = 0;
    for (size_t i = 0; i < rows.seq.size(); i++) {
        const Node &row = rows.seq[i];
        if (row.kind != Node::Seq) throw Error("WorkUnitResult.rows: expected a sequence of sequences");
        if (i == 0) w = (uint32_t)row.seq.size();
        else if (row.seq.size() != w) throw Error("WorkUnitResult.rows: rows differ in length");
        for (const Node &c : row.seq) {
            r.rows.push_back(as_double(struct_field(c, 0, "r", "Color"), "Color.r"));
            r.rows.push_back(as_double(struct_field(c, 1, "g", "Color"), "Color.g"));
            r.rows.push_back(as_double(struct_field(c, 2, "b", "Color"), "Color.b"));
        }
    }
    if (width) *width = w;
    return r;
}

// hands the reader one byte per call, so that `given` is exactly what the first item occupied
struct CountingMemory : cbor::ByteSource {
    const uint8_t *p;
    size_t n, given = 0;
    CountingMemory(const void *d, size_t len) : p(static_cast<const uint8_t *>(d)), n(len) {}
    size_t read(uint8_t *dst, size_t want) override {
        if (given == n || want == 0) return 0;
        dst[0] = p[given++];
        return 1;
    }
};

Node first_item(const void *data, size_t n, size_t *used) {
    CountingMemory src(data, n);
    cbor::Reader rd(src);
    Node node;
    if (!rd.next(node)) throw Error("cbor: empty input");
    if (used) *used = src.given;
    return node;
}

// ---- sockets ----
struct FdSource : cbor::ByteSource {
    int fd;
    explicit FdSource(int f) : fd(f) {}
    size_t read(uint8_t *dst, size_t n) override {
        for (;;) {
            const ssize_t k = ::recv(fd, dst, n, 0);
            if (k >= 0) return (size_t)k;
            if (errno == EINTR) continue;
            throw Error(std::string("recv: ") + std::strerror(errno));
        }
    }
};

void send_all(int fd, const std::string &buf) {
    size_t off = 0;
    while (off < buf.size()) {
        const ssize_t k = ::send(fd, buf.data() + off, buf.size() - off, MSG_NOSIGNAL);
        if (k < 0) {
            if (errno == EINTR) continue;
            throw Error(std::string("send: ") + std::strerror(errno));
        }
        off += (size_t)k;
    }
}

}  // namespace

std::string encode_worker_info(uint64_t num_threads) {
    Writer w;
    w.map(1);
    w.key("num_threads").uint(num_threads);
    return w.out;
}

std::string encode_set_job(const Job &job, EnumForm form) {
    Writer w;
    w.legacy_enums = form == EnumForm::Array;
    w.variant("SetJob");
    w.map(3);
    w.key("id");
    put_job_id(w, job.id.allocator_id, job.id.id);
    w.key("scene_data");
    put_scene(w, job.scene_data);
    w.key("config").map(3);
    w.key("sample_root").uint(job.config.sample_root);
    w.key("max_trace_depth").uint(job.config.max_trace_depth);
    w.key("rows_per_work_unit").uint(job.config.rows_per_work_unit);
    return w.out;
}

std::string encode_work_unit(const WorkUnit &unit, EnumForm form) {
    Writer w;
    w.legacy_enums = form == EnumForm::Array;
    w.variant("WorkUnit");
    put_work_unit(w, unit);
    return w.out;
}

std::string encode_done() {
    Writer w;
    w.text("Done");
    return w.out;
}

std::string encode_rows_ready(const WorkUnitResult &r, uint32_t width, EnumForm form) {
    const size_t row_elems = (size_t)width * 3;
    if (row_elems == 0 || r.rows.size() % row_elems) throw Error("encode_rows_ready: rows do not match the width");
    const size_t n_rows = r.rows.size() / row_elems;
    Writer w;
    w.out.reserve(64 + r.rows.size() * 10);
    w.legacy_enums = form == EnumForm::Array;
    w.variant("RowsReady");
    w.map(2);
    w.key("work_unit");
    put_work_unit(w, r.work_unit);
    w.key("rows").array(n_rows);
    const double *p = r.rows.data();
    for (size_t y = 0; y < n_rows; y++) {
        w.array(width);
        for (uint32_t x = 0; x < width; x++, p += 3) {
            w.out.append("\xa3\x61r", 3);   // map(3), text(1) "r"
            w.f64(p[0]);
            w.out.append("\x61g", 2);
            w.f64(p[1]);
            w.out.append("\x61" "b", 2);
            w.f64(p[2]);
        }
    }
    return w.out;
}

Request decode_request(const void *data, size_t n, size_t *used) { return request_from(first_item(data, n, used)); }
uint64_t decode_worker_info(const void *data, size_t n, size_t *used) { return worker_info_from(first_item(data, n, used)); }
WorkUnitResult decode_rows_ready(const void *data, size_t n, uint32_t *width, size_t *used) {
    return rows_ready_from(first_item(data, n, used), width);
}

// ------------------------------------------------------------------------------------------------
// NodeServer
// ------------------------------------------------------------------------------------------------
NodeServer::NodeServer(GpuWorker &worker, const std::string &host, const std::string &port) : worker_(worker) {
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    hints.ai_flags = AI_PASSIVE | AI_NUMERICSERV;
    const int rc = getaddrinfo(host.empty() ? nullptr : host.c_str(), port.c_str(), &hints, &res);
    if (rc != 0) throw Error("cannot resolve `" + host + ":" + port + "`: " + gai_strerror(rc));
    std::string last = "no address";
    for (addrinfo *a = res; a; a = a->ai_next) {
        const int fd = ::socket(a->ai_family, a->ai_socktype, a->ai_protocol);
        if (fd < 0) {
            last = std::strerror(errno);
            continue;
        }
        const int one = 1;
        setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        if (::bind(fd, a->ai_addr, a->ai_addrlen) == 0 && ::listen(fd, 16) == 0) {
            sockaddr_storage ss{};
            socklen_t len = sizeof ss;
            getsockname(fd, reinterpret_cast<sockaddr *>(&ss), &len);
            port_ = ntohs(ss.ss_family == AF_INET6 ? reinterpret_cast<sockaddr_in6 *>(&ss)->sin6_port
                                                   : reinterpret_cast<sockaddr_in *>(&ss)->sin_port);
            listen_fd_ = fd;
            break;
        }
        last = std::strerror(errno);
        ::close(fd);
    }
    freeaddrinfo(res);
    if (listen_fd_ < 0) throw Error("cannot bind `" + host + ":" + port + "`: " + last);
}

NodeServer::~NodeServer() {
    if (listen_fd_ >= 0) ::close(listen_fd_);
}

void NodeServer::stop() {
    stop_.store(true);
    if (listen_fd_ >= 0) ::shutdown(listen_fd_, SHUT_RDWR);   // wakes accept()
}

void NodeServer::serve(uint64_t max_clients) {
    for (uint64_t served = 0; !stop_.load() && (max_clients == 0 || served < max_clients); served++) {
        sockaddr_storage ss{};
        socklen_t len = sizeof ss;
        const int fd = ::accept(listen_fd_, reinterpret_cast<sockaddr *>(&ss), &len);
        if (fd < 0) {
            if (errno == EINTR) {
                served--;
                continue;
            }
            if (stop_.load()) return;
            throw Error(std::string("accept: ") + std::strerror(errno));
        }
        char hostbuf[NI_MAXHOST] = "?", servbuf[NI_MAXSERV] = "?";
        getnameinfo(reinterpret_cast<sockaddr *>(&ss), len, hostbuf, sizeof hostbuf, servbuf, sizeof servbuf, NI_NUMERICHOST | NI_NUMERICSERV);
        const std::string peer = std::string(hostbuf) + ":" + servbuf;
        try {
            handle_client(fd, peer);
        } catch (const std::exception &e) {
            std::printf("run_server: handle_client exited with %s\n", e.what());   // flux-node/src/main.rs:104-106
            std::fflush(stdout);
        }
        ::close(fd);
    }
}

// flux-node/src/main.rs:21-94.  The result thread of the reference is kept: RowsReady of unit k is encoded and
// sent while unit k+1 renders (the manager keeps two units in flight, workers.rs:160-175).
void NodeServer::handle_client(int fd, const std::string &peer) {
    std::printf("Got connection from %s\n", peer.c_str());
    std::fflush(stdout);
    const int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    send_all(fd, encode_worker_info(worker_.info().num_threads));

    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::pair<WorkUnitResult, uint32_t>> queue;
    std::atomic<EnumForm> form{EnumForm::Array};   // answer in the form the manager writes (serde_cbor < 0.10 reads no other)
    bool closing = false;
    std::string send_error;
    std::thread sender([&] {
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return closing || !queue.empty(); });
            if (queue.empty()) return;
            auto item = std::move(queue.front());
            queue.pop_front();
            lk.unlock();
            try {
                send_all(fd, encode_rows_ready(item.first, item.second, form.load()));
            } catch (const std::exception &e) {
                std::lock_guard<std::mutex> g(mu);
                send_error = e.what();   // "Manager connection error" in the reference: stop sending
                return;
            }
        }
    });
    auto finish = [&] {
        {
            std::lock_guard<std::mutex> g(mu);
            closing = true;
        }
        cv.notify_all();
        sender.join();
    };

    try {
        FdSource src(fd);
        cbor::Reader reader(src);
        Node node;
        while (reader.next(node)) {
            Request req = request_from(node);
            if (req.kind != Request::Done) form.store(req.form);
            if (req.kind == Request::SetJob) {
                std::printf("Got job\n");
                std::fflush(stdout);
                worker_.begin_job(req.job.scene_data, req.job.config);
            } else if (req.kind == Request::Unit) {
                if (!worker_.has_job()) throw Error("work unit before SetJob");
                WorkUnitResult r = worker_.render_unit(req.unit);
                {
                    std::lock_guard<std::mutex> g(mu);
                    if (!send_error.empty()) throw Error("Manager connection error: " + send_error);
                    queue.emplace_back(std::move(r), worker_.job_image_width());
                }
                cv.notify_all();
            } else {
                std::printf("Got done message, shutting down\n");
                std::fflush(stdout);
                break;
            }
        }
    } catch (...) {
        finish();
        throw;
    }
    finish();
    if (!send_error.empty()) throw Error("Manager connection error: " + send_error);
}

// ------------------------------------------------------------------------------------------------
// NetworkWorker
// ------------------------------------------------------------------------------------------------
NetworkWorker::NetworkWorker(const std::string &raw_endpoint, EnumForm form) : form_(form) {
    // workers.rs:119-123: "host" gets the default port
    const size_t colon = raw_endpoint.rfind(':');
    const std::string host = colon == std::string::npos ? raw_endpoint : raw_endpoint.substr(0, colon);
    const std::string port = colon == std::string::npos ? DEFAULT_PORT : raw_endpoint.substr(colon + 1);
    endpoint_ = host + ":" + port;
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    const int rc = getaddrinfo(host.c_str(), port.c_str(), &hints, &res);
    if (rc != 0) throw Error("cannot resolve `" + endpoint_ + "`: " + gai_strerror(rc));
    std::string last = "no address";
    for (addrinfo *a = res; a && fd_ < 0; a = a->ai_next) {
        const int fd = ::socket(a->ai_family, a->ai_socktype, a->ai_protocol);
        if (fd < 0) continue;
        if (::connect(fd, a->ai_addr, a->ai_addrlen) == 0) fd_ = fd;
        else {
            last = std::strerror(errno);
            ::close(fd);
        }
    }
    freeaddrinfo(res);
    if (fd_ < 0) throw Error("cannot connect to `" + endpoint_ + "`: " + last);
    const int one = 1;
    setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
}

NetworkWorker::~NetworkWorker() {
    if (fd_ >= 0) ::close(fd_);
}

bool UnitQueue::pop(WorkUnit &u) {
    std::lock_guard<std::mutex> g(mu_);
    if (next_ >= units_.size()) return false;
    u = units_[next_++];
    return true;
}

UnitQueue::UnitQueue(const Job &job) : units_(work_units(job.scene_data.output_settings.image_height, job.config.rows_per_work_unit, job.id.id)) {
    for (WorkUnit &u : units_) u.job_allocator_id = job.id.allocator_id;
}

void NetworkWorker::run(const Job &job, UnitQueue &queue, Image &img) {
    if (fd_ < 0) throw Error("NetworkWorker: connection already used (the node ends it after Done)");
    FdSource src(fd_);
    cbor::Reader reader(src);
    Node node;
    // workers.rs:137-143: the first thing on the stream is the node's WorkerInfo
    if (!reader.next(node)) throw Error("Could not get info from network node");
    info_ = WorkerInfo{"NetworkWorker(" + endpoint_ + ")", (uint32_t)worker_info_from(node)};

    const uint32_t W = job.scene_data.output_settings.image_width;
    send_all(fd_, encode_set_job(job, form_));
    auto collect = [&] {
        if (!reader.next(node)) throw Error("network node closed the connection before all results arrived");
        uint32_t w = 0;
        WorkUnitResult r = rows_ready_from(node, &w);
        if (w != W) throw Error("network node returned rows of another width");
        if (r.work_unit.job_id != job.id.id || r.work_unit.job_allocator_id != job.id.allocator_id)
            throw Error("network node returned rows of another job");
        img.set_rows(r);   // units are disjoint row ranges: workers sharing an Image never touch the same bytes
    };
    // two units in flight (workers.rs:160-175), then one more unit per result while the shared queue has any
    // (the reference's workers pull from one bounded queue the same way, manager.rs:100,156-162), then the tail
    size_t in_flight = 0;
    WorkUnit u{0, 0, 0, 0};
    while (in_flight < 2 && queue.pop(u)) {
        send_all(fd_, encode_work_unit(u, form_));
        in_flight++;
    }
    while (in_flight) {
        if (queue.pop(u)) {
            send_all(fd_, encode_work_unit(u, form_));
            in_flight++;
        }
        collect();
        in_flight--;
    }
    send_all(fd_, encode_done());
    ::close(fd_);
    fd_ = -1;
}

Image NetworkWorker::render_job(const Job &job) {
    Image img(job.scene_data.output_settings.image_width, job.scene_data.output_settings.image_height);
    UnitQueue queue(job);
    run(job, queue, img);
    return img;
}

Image render_job_on_nodes(const std::vector<std::string> &endpoints, const Job &job, EnumForm form, std::vector<WorkerInfo> *infos,
                          GpuWorker *local) {
    if (endpoints.empty() && !local) throw Error("render_job_on_nodes: no workers");
    Image img(job.scene_data.output_settings.image_width, job.scene_data.output_settings.image_height);
    UnitQueue queue(job);
    std::vector<std::unique_ptr<NetworkWorker>> workers;
    for (const std::string &e : endpoints) workers.push_back(std::make_unique<NetworkWorker>(e, form));   // connect first: fail early
    std::vector<std::string> errors(workers.size() + 1);
    std::vector<std::thread> th;
    if (local)
        th.emplace_back([&] {
            try {   // the LocalWorker loop, workers.rs:46-71
                local->begin_job(job.scene_data, job.config);
                WorkUnit u{0, 0, 0, 0};
                while (queue.pop(u)) img.set_rows(local->render_unit(u));
            } catch (const std::exception &e) {
                errors[workers.size()] = e.what();
            }
        });
    for (size_t i = 0; i < workers.size(); i++)
        th.emplace_back([&, i] {
            try {
                workers[i]->run(job, queue, img);
            } catch (const std::exception &e) {
                errors[i] = e.what();   // like the reference, a lost node is fatal for the job (manager.rs:158-161)
            }
        });
    for (auto &t : th) t.join();
    for (const std::string &e : errors)
        if (!e.empty()) throw Error(e);
    if (infos) {
        if (local) infos->push_back(local->info());
        for (auto &w : workers) infos->push_back(w->info());
    }
    return img;
}

}  // namespace net
}  // namespace flux
