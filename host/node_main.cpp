// fluxb200-node — network rendering server in place of the reference's `flux-node` (flux-node/src/main.rs:113-168),
// rendering on the GPUs of this box:
//
//   fluxb200-node [-h ADDRESS] [-p PORT] [-G GPUS] [--seed S] [--clients N]
//
// -h/--host and -p/--port keep the reference's meaning and defaults (0.0.0.0, 2000).  -t/--threads (CPU render
// threads) has no counterpart and is accepted and ignored; WorkerInfo.num_threads reports the GPU count.
// --clients N makes the server exit after N connections (tests).  An unmodified manager connects with
// `flux scenes/demo2.yml -L -n host:port`.
//
// Codec tools (no GPU needed; used by the CPU tests):
//   --decode FILE        print one line per NetworkWorkerRequest found in FILE
//   --decode-flat FILE OUT   decode the SetJob at the front of FILE and write its flattened scene to OUT
//   --reencode FILE      decode the requests in FILE and write them again to stdout
//   --rows-ready ROW_START ROW_END WIDTH ALLOC_ID JOB_ID   read raw f64 RGB from stdin, write RenderEvent::RowsReady
//   --enum-form array|map   (before --rows-ready) the enum form to write: serde_cbor < 0.10 (default) or >= 0.10
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>

#include "fluxnet.hpp"

namespace {

[[noreturn]] void usage(const char *msg) {
    if (msg) std::fprintf(stderr, "error: %s\n\n", msg);
    std::fprintf(stderr,
                 "fluxb200-node — network rendering server for the flux ray tracer, B200 render path\n\n"
                 "USAGE:\n    fluxb200-node [OPTIONS]\n\nOPTIONS:\n"
                 "    -h, --host <ADDRESS>   Listen for requests on this address [default: 0.0.0.0]\n"
                 "    -p, --port <PORT>      Listen on this TCP port [default: 2000]\n"
                 "    -G, --gpus <N>         GPUs of this box to render on [default: 1]\n"
                 "        --devices <LIST>   Explicit CUDA device list instead of -G, e.g. 0,2,3\n"
                 "        --seed <S>         Seed of the sample sets [default: 1]\n"
                 "        --clients <N>      Exit after serving N connections [default: serve forever]\n"
                 "        --decode <FILE> | --reencode <FILE> | --rows-ready <5 integers>   codec tools, see the source\n");
    std::exit(msg ? 2 : 0);
}

uint64_t parse_u64(const char *s, const char *what) {
    char *end = nullptr;
    if (!s || !*s || *s == '-') usage((std::string("invalid value for ") + what).c_str());
    const unsigned long long v = std::strtoull(s, &end, 10);
    if (*end) usage((std::string("invalid value for ") + what).c_str());
    return v;
}

std::string slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw flux::Error("cannot read `" + path + "`");
    return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int codec_requests(const std::string &path, bool reencode) {
    const std::string data = slurp(path);
    size_t off = 0;
    while (off < data.size()) {
        size_t used = 0;
        const flux::net::Request r = flux::net::decode_request(data.data() + off, data.size() - off, &used);
        off += used;
        if (reencode) {
            const std::string out = r.kind == flux::net::Request::SetJob ? flux::net::encode_set_job(r.job, r.form)
                                    : r.kind == flux::net::Request::Unit ? flux::net::encode_work_unit(r.unit, r.form)
                                                                         : flux::net::encode_done();
            std::fwrite(out.data(), 1, out.size(), stdout);
        } else if (r.kind == flux::net::Request::SetJob) {
            const flux::SceneData &s = r.job.scene_data;
            auto flat = s.flatten();
            std::printf("SetJob id=(%llu,%llu) scene=%s image=%ux%u shapes=%zu spheres=%u planes=%u triangles=%u materials=%u "
                        "root=%u depth=%u rows=%u eye=%a,%a,%a lens_radius=%a\n",
                        (unsigned long long)r.job.id.allocator_id, (unsigned long long)r.job.id.id, s.scene_name.c_str(),
                        s.output_settings.image_width, s.output_settings.image_height, s.shapes.size(), flat->flat.n_spheres,
                        flat->flat.n_planes, flat->flat.n_triangles, flat->flat.n_materials, r.job.config.sample_root,
                        r.job.config.max_trace_depth, r.job.config.rows_per_work_unit, s.camera_settings.eye[0], s.camera_settings.eye[1],
                        s.camera_settings.eye[2], s.camera_data.lens_radius);
        } else if (r.kind == flux::net::Request::Unit) {
            std::printf("WorkUnit rows=%u..%u job=(%llu,%llu)\n", r.unit.row_start, r.unit.row_end,
                        (unsigned long long)r.unit.job_allocator_id, (unsigned long long)r.unit.job_id);
        } else {
            std::printf("Done\n");
        }
    }
    return 0;
}

flux::net::EnumForm g_form = flux::net::EnumForm::Array;

int codec_rows_ready(char **v) {
    flux::WorkUnit u{(uint32_t)parse_u64(v[0], "ROW_START"), (uint32_t)parse_u64(v[1], "ROW_END"), parse_u64(v[4], "JOB_ID"),
                     parse_u64(v[3], "ALLOC_ID")};
    const uint32_t width = (uint32_t)parse_u64(v[2], "WIDTH");
    const std::string raw((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
    if (raw.size() % sizeof(double)) throw flux::Error("stdin is not a whole number of doubles");
    flux::WorkUnitResult r{u, std::vector<double>(raw.size() / sizeof(double))};
    std::memcpy(r.rows.data(), raw.data(), raw.size());
    const std::string out = flux::net::encode_rows_ready(r, width, g_form);
    std::fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    std::string host = "0.0.0.0", port = flux::net::DEFAULT_PORT;
    uint32_t gpus = 1;
    uint64_t seed = 1, clients = 0;
    std::vector<int> devices;
    try {
        for (int i = 1; i < argc; i++) {
            const std::string a = argv[i];
            auto next = [&](const char *what) -> const char * {
                if (i + 1 >= argc) usage((std::string("missing value for ") + what).c_str());
                return argv[++i];
            };
            if (a == "-h" || a == "--host") host = next("--host");
            else if (a == "-p" || a == "--port") port = next("--port");
            else if (a == "-t" || a == "--threads") next("--threads");
            else if (a == "-G" || a == "--gpus") gpus = (uint32_t)parse_u64(next("--gpus"), "--gpus");
            else if (a == "--devices") {   // explicit device list, e.g. 0,0 = two contexts on one GPU (tests of the sharded path)
                for (const char *p = next("--devices"); *p;) {
                    char *end = nullptr;
                    devices.push_back((int)std::strtol(p, &end, 10));
                    if (end == p) usage("invalid value for --devices");
                    p = *end == ',' ? end + 1 : end;
                }
            } else if (a == "--seed") seed = parse_u64(next("--seed"), "--seed");
            else if (a == "--clients") clients = parse_u64(next("--clients"), "--clients");
            else if (a == "--enum-form") {   // codec tools only: the server mirrors its client
                const std::string f = next("--enum-form");
                if (f != "array" && f != "map") usage("--enum-form takes array or map");
                g_form = f == "map" ? flux::net::EnumForm::Map : flux::net::EnumForm::Array;
            }
            else if (a == "--decode") return codec_requests(next("--decode"), false);
            else if (a == "--reencode") return codec_requests(next("--reencode"), true);
            else if (a == "--decode-flat") {
                const std::string data = slurp(next("--decode-flat"));
                const flux::net::Request r = flux::net::decode_request(data.data(), data.size());
                if (r.kind != flux::net::Request::SetJob) throw flux::Error("--decode-flat: the first request is not SetJob");
                auto flat = r.job.scene_data.flatten();
                flux::dump_flat_text(next("OUT"), *flat);
                return 0;
            }
            else if (a == "--rows-ready") {
                if (i + 5 >= argc) usage("--rows-ready takes 5 integers");
                return codec_rows_ready(argv + i + 1);
            } else if (a == "--help") usage(nullptr);
            else usage(("unknown option " + a).c_str());
        }
        if (gpus == 0) usage("--gpus must be >= 1");
        if (devices.empty())
            for (uint32_t g = 0; g < gpus; g++) devices.push_back((int)g);
        flux::GpuWorker worker(devices, seed);   // fails loudly without a GPU: there is no CPU rendering path
        flux::net::NodeServer server(worker, host, port);
        std::printf("Bind address: %s:%u\n", host.c_str(), (unsigned)server.port());   // flux-node/src/main.rs:159
        std::fflush(stdout);
        server.serve(clients);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "fluxb200-node: %s\n", e.what());
        return 101;
    }
    return 0;
}
