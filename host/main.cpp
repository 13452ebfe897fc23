// fluxb200 — render driver mirroring the reference's `flux` binary (flux/src/main.rs:23-205) for the GPU path:
//
//   fluxb200 <scene_file> [-r ROOT] [-d DEPTH] [-R COUNT] [-G GPUS] [--seed S] [--width W --height H] [-o FILE]
//            [--progressive K] [-n ADDRESS[:PORT]]
//
// -r/--root, -d/--depth and -R/--rows keep the reference's meaning and defaults (1, 5, 50).  -n/--node ADDRESS[:PORT]
// renders on fluxb200-node (or flux-node) processes instead of the local GPUs, speaking the reference's network
// protocol (workers.rs:118-245); given several times, the nodes share the job's work units through one queue, and
// so do the GPUs of this box unless -L is given (the reference's local worker, main.rs:43-60).  -L, -g and -t have no
// counterpart: there is no local CPU worker and no SDL preview.  The image is written to <scene_name>.ppm like ImageBuilder does
// (manager.rs:330) unless -o is given.  --dump-flat FILE writes the flattened scene (no GPU needed; used by the
// tests to compare this loader with the Python mirror).
#include <chrono>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "fluxhost.hpp"
#include "fluxnet.hpp"

namespace {

volatile std::sig_atomic_t g_interrupted = 0;

struct Config {   // flux/src/main.rs:114-124
    std::string input_filename, output_filename, dump_flat;
    std::vector<std::string> nodes;   // -n may be given several times (flux/src/main.rs:131-137)
    uint32_t sample_root = 1, max_depth = 5, rows_per_work_unit = 50;
    uint32_t gpus = 1, width = 0, height = 0;
    uint64_t seed = 1;
    std::vector<int> devices;
    bool enum_map = false, skip_local = false;
    uint32_t progressive = 0;
};

[[noreturn]] void usage(const char *msg) {
    if (msg) std::fprintf(stderr, "error: %s\n\n", msg);
    std::fprintf(stderr,
                 "fluxb200 — flux ray tracer, B200 render path\n\n"
                 "USAGE:\n    fluxb200 [OPTIONS] <scene_file>\n\nOPTIONS:\n"
                 "    -r, --root <ROOT>      Sample root (samples per pixel = ROOT^2) [default: 1]\n"
                 "    -d, --depth <DEPTH>    Tracing depth [default: 5]\n"
                 "    -R, --rows <COUNT>     Image rows per work unit [default: 50]\n"
                 "    -G, --gpus <N>         GPUs of this box to render on [default: 1]\n"
                 "        --devices <LIST>   Explicit CUDA device list instead of -G, e.g. 0,2,3\n"
                 "        --progressive <K>  Refine the frame in passes of K samples per pixel; Ctrl-C keeps what is done\n"
                 "    -L                     Do not use the GPUs of this box for rendering (with -n)\n"
                 "    -n, --node <ADDRESS[:PORT]>   Render using the fluxb200-node / flux-node process at this address\n"
                 "        --enum-form <array|map>   With -n: CBOR form of enum variants, serde_cbor < 0.10 (default) or >= 0.10\n"
                 "        --seed <S>         Seed of the sample sets [default: 1]\n"
                 "        --width <W> --height <H>   Override the scene's image size\n"
                 "    -o <FILE>              Output file [default: <scene_name>.ppm]\n"
                 "        --dump-flat <FILE> Write the flattened scene and exit (no GPU needed)\n");
    std::exit(msg ? 2 : 0);
}

uint64_t parse_u64(const char *s, const char *what) {
    char *end = nullptr;
    if (!s || !*s || *s == '-') usage((std::string("invalid value for ") + what).c_str());
    const unsigned long long v = std::strtoull(s, &end, 10);
    if (*end) usage((std::string("invalid value for ") + what).c_str());   // usize::from_str(..).unwrap() panics
    return v;
}

Config config_from_args(int argc, char **argv) {
    Config c;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&](const char *what) -> const char * {
            if (i + 1 >= argc) usage((std::string("missing value for ") + what).c_str());
            return argv[++i];
        };
        if (a == "-r" || a == "--root") c.sample_root = (uint32_t)parse_u64(next("--root"), "--root");
        else if (a == "-d" || a == "--depth") c.max_depth = (uint32_t)parse_u64(next("--depth"), "--depth");
        else if (a == "-R" || a == "--rows") c.rows_per_work_unit = (uint32_t)parse_u64(next("--rows"), "--rows");
        else if (a == "-G" || a == "--gpus") c.gpus = (uint32_t)parse_u64(next("--gpus"), "--gpus");
        else if (a == "-L") c.skip_local = true;   // flux/src/main.rs:150-153
        else if (a == "--progressive") c.progressive = (uint32_t)parse_u64(next("--progressive"), "--progressive");
        else if (a == "-n" || a == "--node") c.nodes.push_back(next("--node"));
        else if (a == "--enum-form") {   // with -n: how enum variants are written (see host/fluxnet.hpp)
            const std::string f = next("--enum-form");
            if (f != "array" && f != "map") usage("--enum-form takes array or map");
            c.enum_map = f == "map";
        }
        else if (a == "--devices") {   // explicit device list, e.g. 0,0 = two contexts on one GPU (tests of the sharded path)
            for (const char *p = next("--devices"); *p;) {
                char *end = nullptr;
                c.devices.push_back((int)std::strtol(p, &end, 10));
                if (end == p) usage("invalid value for --devices");
                p = *end == ',' ? end + 1 : end;
            }
        } else if (a == "--seed") c.seed = parse_u64(next("--seed"), "--seed");
        else if (a == "--width") c.width = (uint32_t)parse_u64(next("--width"), "--width");
        else if (a == "--height") c.height = (uint32_t)parse_u64(next("--height"), "--height");
        else if (a == "-o") c.output_filename = next("-o");
        else if (a == "--dump-flat") c.dump_flat = next("--dump-flat");
        else if (a == "-h" || a == "--help") usage(nullptr);
        else if (!a.empty() && a[0] == '-') usage(("unknown option " + a).c_str());
        else if (c.input_filename.empty()) c.input_filename = a;
        else usage("more than one scene file");
    }
    if (c.input_filename.empty()) usage("Scene filename is required");
    if (c.sample_root == 0 || c.gpus == 0 || c.rows_per_work_unit == 0) usage("--root, --gpus and --rows must be >= 1");
    return c;
}

}  // namespace

int main(int argc, char **argv) {
    const Config config = config_from_args(argc, argv);
    try {
        // Load the YAML scene file (flux/src/main.rs:27-29)
        flux::SceneData s = flux::SceneData::from_yaml_file(config.input_filename);
        if (config.width && config.height) s = s.with_size(config.width, config.height);
        if (!config.dump_flat.empty()) {
            auto flat = s.flatten();
            flux::dump_flat_text(config.dump_flat, *flat);
            return 0;
        }
        flux::JobConfiguration netcfg;
        netcfg.sample_root = config.sample_root;
        netcfg.max_trace_depth = config.max_depth;
        netcfg.rows_per_work_unit = config.rows_per_work_unit;
        if (!config.nodes.empty()) {
            std::printf("flux render (%s, %u sample%s per pixel, max depth %u)\n", s.scene_name.c_str(),
                        netcfg.sample_root * netcfg.sample_root, netcfg.sample_root == 1 ? "" : "s", netcfg.max_trace_depth);
            std::vector<flux::WorkerInfo> infos;
            std::unique_ptr<flux::GpuWorker> local;   // the local worker renders beside the nodes unless -L (main.rs:43-60)
            if (!config.skip_local) {
                std::vector<int> devices = config.devices;
                if (devices.empty())
                    for (uint32_t g = 0; g < config.gpus; g++) devices.push_back((int)g);
                local = std::make_unique<flux::GpuWorker>(devices, config.seed);
            }
            flux::Image img = flux::net::render_job_on_nodes(config.nodes, flux::Job{flux::JobID{config.seed, 0}, s, netcfg},
                                                             config.enum_map ? flux::net::EnumForm::Map : flux::net::EnumForm::Array, &infos,
                                                             local.get());
            for (const flux::WorkerInfo &wi : infos) std::printf("%s ready, info:\nThreads: %u\n", wi.name.c_str(), wi.num_threads);
            const std::string out = config.output_filename.empty() ? s.scene_name + ".ppm" : config.output_filename;
            img.write(out);
            std::printf("wrote %s\nShutting down\n", out.c_str());
            return 0;
        }
        std::vector<int> devices = config.devices;
        if (devices.empty())
            for (uint32_t g = 0; g < config.gpus; g++) devices.push_back((int)g);
        flux::GpuWorker worker(devices, config.seed);
        std::printf("GPU worker ready, info:\n");
        std::printf("Threads: %u\n", worker.info().num_threads);   // WorkerInfo::print, manager.rs:227-229
        flux::JobConfiguration jobcfg;
        jobcfg.sample_root = config.sample_root;
        jobcfg.max_trace_depth = config.max_depth;
        jobcfg.rows_per_work_unit = config.rows_per_work_unit;
        std::printf("flux render (%s, %u sample%s per pixel, max depth %u)\n", s.scene_name.c_str(),   // title(), main.rs:207-214
                    jobcfg.sample_root * jobcfg.sample_root, jobcfg.sample_root == 1 ? "" : "s", jobcfg.max_trace_depth);
        if (config.progressive) {
            // the preview's refinement and Esc (flux/src/main.rs:288-315) without the window: every pass improves
            // the whole frame; an interrupt stops after the pass in flight and the image so far is written
            std::signal(SIGINT, [](int) { g_interrupted = 1; });
            const auto t0 = std::chrono::steady_clock::now();
            uint32_t samples = 0;
            flux::Image img = worker.render_job_progressive(s, jobcfg, config.progressive, [&](uint32_t done, const flux::Image &) {
                samples = done;
                std::printf("pass to %u samples per pixel, %.3fs\n", done,
                            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
                std::fflush(stdout);
                return !g_interrupted;
            });
            const uint32_t total = jobcfg.sample_root * jobcfg.sample_root;
            if (samples < total) std::printf("cancelled after %u of %u samples per pixel\n", samples, total);
            else std::printf("rendering finished, total time %.6fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            const std::string out = config.output_filename.empty() ? s.scene_name + ".ppm" : config.output_filename;
            img.write(out);
            std::printf("wrote %s\nShutting down\n", out.c_str());
            return 0;
        }
        double seconds = 0.0;
        flux::Image img = worker.render_job(s, jobcfg, &seconds);
        std::printf("rendering finished, total time %.6fs\n", seconds);   // manager.rs:327
        const double n = (double)img.width * img.height * jobcfg.sample_root * jobcfg.sample_root;
        std::printf("%.1f Msamples/s on %u GPU(s)\n", n / seconds / 1e6, (unsigned)devices.size());
        const std::string out = config.output_filename.empty() ? s.scene_name + ".ppm" : config.output_filename;
        img.write(out);
        std::printf("wrote %s\n", out.c_str());
        std::printf("Shutting down\n");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "fluxb200: %s\n", e.what());   // the reference unwrap()s: a panic with the error text
        return 101;                                          // Rust's panic exit status
    }
    return 0;
}
