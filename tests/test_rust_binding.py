"""The Rust binding (rust/fluxb200-sys, rust/gpu_worker.rs) cannot be compiled here — the image has no cargo / rustc —
so it is held to the C header textually: every function of include/fluxb200.h is declared in src/lib.rs with the same
number of parameters, the #[repr(C)] structs list the header's fields in the header's order with matching widths, and
the worker only calls functions the binding declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_c(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def _header():
    return _strip_c(open(os.path.join(ROOT, "include", "fluxb200.h")).read())


def _rust():
    text = open(os.path.join(ROOT, "rust", "fluxb200-sys", "src", "lib.rs")).read()
    return re.sub(r"//[^\n]*", "", text)


def _count_params(arglist):
    arglist = arglist.strip()
    if arglist in ("", "void"):
        return 0
    return len([a for a in arglist.split(",") if a.strip()])


def c_functions():
    out = {}
    for m in re.finditer(r"\b(flux_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", _header()):
        out[m.group(1)] = _count_params(m.group(2))
    return out


def rust_functions():
    out = {}
    for m in re.finditer(r"pub fn (flux_[a-z0-9_]+)\s*\(([^()]*)\)", _rust(), flags=re.S):
        out[m.group(1)] = _count_params(m.group(2))
    return out


def test_every_header_function_is_bound_with_the_same_arity():
    c, r = c_functions(), rust_functions()
    assert len(c) >= 40
    assert sorted(c) == sorted(r)
    assert c == r


C2RUST = {"uint32_t": "u32", "uint64_t": "u64", "uint8_t": "u8", "int32_t": "i32", "double": "f64", "float": "f32"}


def c_struct_fields(name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), _header(), flags=re.S)
    assert m, name
    fields = []
    for decl in m.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m2 = re.match(r"(const )?([a-z0-9_]+) (\*?)([a-zA-Z0-9_]+)(\[(\d+)\])?$", decl)
        assert m2, decl
        const, ctype, ptr, fname, _, arr = m2.groups()
        rtype = C2RUST.get(ctype, ctype)
        if ptr:
            rtype = "*const " + rtype
        if arr:
            rtype = "[%s; %s]" % (rtype, arr)
        fields.append((fname, rtype))
    return fields


def rust_struct_fields(name):
    m = re.search(r"pub struct %s \{(.*?)\n\}" % name, _rust(), flags=re.S)
    assert m, name
    return [(f.strip(), t.strip()) for f, t in re.findall(r"pub ([a-zA-Z0-9_]+):\s*([^,\n]+),", m.group(1))]


def test_repr_c_structs_follow_the_header_field_for_field():
    for name in ("flux_material", "flux_scene_flat", "flux_job_config", "flux_counters"):
        assert rust_struct_fields(name) == c_struct_fields(name), name
    assert "#[repr(C)]\n#[derive(Clone, Copy, Debug)]\npub struct flux_material" in open(
        os.path.join(ROOT, "rust", "fluxb200-sys", "src", "lib.rs")).read()


def test_constants_match_the_header():
    h, r = _header(), _rust()
    for name in ("FLUX_OK", "FLUX_ERR_INVALID", "FLUX_ERR_CUDA", "FLUX_ERR_STATE", "FLUX_ERR_NO_DEVICE",
                 "FLUX_MAT_MATTE", "FLUX_MAT_EMISSIVE", "FLUX_MAT_REFLECTIVE", "FLUX_MAT_GLOSSY"):
        hv = re.search(r"\b%s = (\d+)" % name, h).group(1)
        rv = re.search(r"pub const %s: [a-z_0-9]+ = (\d+);" % name, r).group(1)
        assert hv == rv, name
    assert re.search(r"#define FLUX_FRAME_HANDLE_BYTES (\d+)", h).group(1) == re.search(
        r"FLUX_FRAME_HANDLE_BYTES: usize = (\d+);", r).group(1)


def test_worker_calls_only_bound_functions_and_implements_the_trait():
    src = open(os.path.join(ROOT, "rust", "gpu_worker.rs")).read()
    code = re.sub(r"//[^\n]*", "", src)
    called = set(re.findall(r"\b(flux_[a-z0-9_]+)\s*\(", code))
    assert called and called <= set(rust_functions())
    assert {"flux_ctx_create", "flux_set_scene", "flux_generate_samples", "flux_render_rows", "flux_ctx_destroy"} <= called
    assert "impl Worker for GpuWorker" in code
    for method in ("fn handle(&self) -> WorkerHandle", "fn stop(self)", "fn info(&self) -> WorkerInfo"):   # manager.rs:232-236
        assert method in code
    assert "UNCOMPILED" in src and "UNCOMPILED" in open(os.path.join(ROOT, "rust", "fluxb200-sys", "src", "lib.rs")).read()
    for f in ("Cargo.toml", "build.rs"):
        assert os.path.exists(os.path.join(ROOT, "rust", "fluxb200-sys", f))
