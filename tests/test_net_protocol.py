"""The network-rendering protocol: wire forms of NetworkWorkerRequest / WorkerInfo /
RenderEvent::RowsReady (fluxcore/src/workers.rs:106-110, manager.rs:16-28,221-224, job.rs:10,40-62).

Two implementations written independently — C++ (host/cbor.cpp, host/fluxnet.cpp, driven through the codec tools of
`fluxb200-node`) and Python (flux_b200/netproto.py) — against each other and against byte vectors derived by hand
from RFC 7049.  No GPU: nothing here renders.  The Rust reference cannot be built in this image, so
these vectors pin the documented serde_cbor forms, not the reference binary.  Enum variants are exercised in both of
serde_cbor's forms: the 2-array of releases < 0.10 (the reference pins 0.9.0; the default here) and the one-entry map
of later releases."""
import os
import struct
import subprocess

import numpy as np
import pytest

from flux_b200 import JobConfiguration, SceneData, WorkUnit
from flux_b200 import netproto as N
from tests import helpers as Hp
from tests.test_cpp_host import _check_same, _parse_dump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "fluxb200-node")


@pytest.fixture(scope="module")
def node():
    subprocess.run(["make", "-C", os.path.join(ROOT, "host"), "-s"], check=True)
    assert os.path.exists(BIN)
    return BIN


def run(node, *args, stdin=None, ok=True):
    p = subprocess.run([node, *args], input=stdin, capture_output=True, timeout=120)
    if ok:
        assert p.returncode == 0, p.stderr.decode()
    return p


# ---- hand-derived vectors (RFC 7049 §2.1: major type << 5 | length; text = 0x60+len; map = 0xa0+n; array = 0x80+n)
WORKER_INFO_8 = bytes.fromhex("a1" "6b") + b"num_threads" + bytes.fromhex("08")
DONE = bytes.fromhex("64") + b"Done"
_WORK_UNIT_BODY = (bytes.fromhex("a3")
                   + bytes.fromhex("69") + b"row_start" + bytes.fromhex("00")
                   + bytes.fromhex("67") + b"row_end" + bytes.fromhex("1831")                 # 49 -> 0x18 0x31
                   + bytes.fromhex("66") + b"job_id" + bytes.fromhex("82" "1b" "0123456789abcdef" "19" "0100"))   # (u64, 256)
# NetworkWorkerRequest::WorkUnit(..): the 2-array [name, content] of serde_cbor < 0.10 (the reference pins 0.9.0) ...
WORK_UNIT = bytes.fromhex("82" "68") + b"WorkUnit" + _WORK_UNIT_BODY
# ... and the one-entry map {name: content} of serde_cbor >= 0.10
WORK_UNIT_MAP = bytes.fromhex("a1" "68") + b"WorkUnit" + _WORK_UNIT_BODY


def test_python_codec_matches_hand_derived_vectors():
    assert N.worker_info(8) == WORKER_INFO_8
    assert N.done() == DONE
    assert N.work_unit(WorkUnit(0, 49, (0x0123456789ABCDEF, 256))) == WORK_UNIT
    assert N.work_unit(WorkUnit(0, 49, (0x0123456789ABCDEF, 256)), "map") == WORK_UNIT_MAP
    # RFC 7049 appendix A examples
    assert N.dumps(1000000) == bytes.fromhex("1a000f4240")
    assert N.dumps(1.5) == bytes.fromhex("fa3fc00000")             # serde_cbor narrows to f32, not to f16
    assert N.dumps(1.1) == bytes.fromhex("fb3ff199999999999a")
    assert N.dumps(100000.0) == bytes.fromhex("fa47c35000")
    assert N.dumps(float("inf")) == bytes.fromhex("f97c00")
    assert N.dumps(float("nan")) == bytes.fromhex("f97e00")
    assert N.dumps([1, [2, 3], [4, 5]]) == bytes.fromhex("8301820203820405")
    assert N.dumps({"a": 1, "b": [2, 3]}) == bytes.fromhex("a26161016162820203")
    for hx, val in (("f93c00", 1.0), ("f9c400", -4.0), ("f90001", 5.960464477539063e-08), ("f97bff", 65504.0),
                    ("fa7f7fffff", 3.4028234663852886e+38), ("fbc010666666666666", -4.1), ("3903e7", -1000),
                    ("9f018202039f0405ffff", [1, [2, 3], [4, 5]]), ("bf6346756ef563416d7421ff", {"Fun": True, "Amt": -2}),
                    ("7f657374726561646d696e67ff", "streaming"), ("c074323031332d30332d32315432303a30343a30305a", "2013-03-21T20:04:00Z")):
        assert N.loads(bytes.fromhex(hx))[0] == val, hx


def test_cpp_decodes_hand_derived_requests(node, tmp_path):
    f = tmp_path / "reqs.cbor"
    f.write_bytes(WORK_UNIT + WORK_UNIT_MAP + DONE)
    out = run(node, "--decode", str(f)).stdout.decode().splitlines()
    assert out == [f"WorkUnit rows=0..49 job=({0x0123456789ABCDEF},256)"] * 2 + ["Done"]
    # shortest-form integers, same key order, each request in the enum form it arrived in
    assert run(node, "--reencode", str(f)).stdout == WORK_UNIT + WORK_UNIT_MAP + DONE


def _scenes():
    yield "demo1", SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo1.yml"))
    yield "demo2", SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo2.yml"))
    yield "deterministic", Hp.deterministic_scene()
    from flux_b200 import synth
    yield "mesh", synth.mesh_scene(12, 7, seed=3, width=40, height=30)      # Mesh extension + env sphere
    yield "glossy", synth.glossy_scene()                                    # all four material kinds


@pytest.mark.parametrize("form", ["array", "map"])
@pytest.mark.parametrize("name,sd", list(_scenes()), ids=[n for n, _ in _scenes()])
def test_set_job_round_trip_between_the_two_codecs(node, tmp_path, name, sd, form):
    """Python encodes SetJob; C++ decodes it into the same flattened scene the Python host hands the C-ABI, and
    encodes it again to the very same bytes."""
    cfg = JobConfiguration(7, 4, 13)
    msg = N.set_job((2 ** 63 + 5, 3), sd, cfg, form)
    assert msg[:1] == (b"\x82" if form == "array" else b"\xa1") and msg[1:8] == b"\x66SetJob"
    f = tmp_path / "job.cbor"
    f.write_bytes(msg + N.work_unit(WorkUnit(13, 25, (2 ** 63 + 5, 3)), form) + N.done())
    lines = run(node, "--decode", str(f)).stdout.decode().splitlines()
    assert len(lines) == 3 and lines[2] == "Done"
    assert lines[0].startswith(f"SetJob id=({2 ** 63 + 5},3) scene={sd.scene_name} "
                               f"image={sd.output_settings.image_width}x{sd.output_settings.image_height} shapes={len(sd.shapes)} ")
    assert " root=7 depth=4 rows=13 " in lines[0]
    assert lines[1] == f"WorkUnit rows=13..25 job=({2 ** 63 + 5},3)"
    dump = tmp_path / "flat.txt"
    run(node, "--decode-flat", str(f), str(dump))
    _check_same(_parse_dump(str(dump)), sd)
    assert run(node, "--reencode", str(f)).stdout == f.read_bytes()


def test_decoder_accepts_every_form_serde_would(node, tmp_path):
    """Indefinite-length containers, f16 / f64 floats where f32 would do, integers for f64 fields, unknown keys,
    Color as a sequence, structs as sequences, the older array form of enums, tags."""
    sd = Hp.deterministic_scene(40, 30)
    tree = {"id": [1, 2], "scene_data": N.scene_tree(sd, "map"), "config": {"sample_root": 2, "max_trace_depth": 5, "rows_per_work_unit": 50}}
    tree["scene_data"]["background"] = [0, 0, 0]                       # Color via visit_seq, integers for f64
    tree["scene_data"]["extra_key_from_a_newer_manager"] = {"x": [1, 2, 3]}
    tree["config"] = [2, 5, 50]                                        # struct via visit_seq
    canonical = N.dumps({"SetJob": tree})

    def indefinite(x):   # every container indefinite-length, every float as f64, text in two chunks
        if isinstance(x, dict):
            return b"\xbf" + b"".join(N.dumps(k) + indefinite(v) for k, v in x.items()) + b"\xff"
        if isinstance(x, list):
            return b"\x9f" + b"".join(indefinite(e) for e in x) + b"\xff"
        if isinstance(x, float):
            return b"\xfb" + struct.pack(">d", x)
        if isinstance(x, str) and len(x) > 2:
            return b"\x7f" + N.dumps(x[:2]) + N.dumps(x[2:]) + b"\xff"
        return N.dumps(x)

    legacy = b"\x82" + N.dumps("SetJob") + b"\xc1" + indefinite(tree)    # ["SetJob", tag(1) content], map-form enums inside
    f1, f2 = tmp_path / "a.cbor", tmp_path / "b.cbor"
    f1.write_bytes(canonical)
    f2.write_bytes(legacy + b"\x82" + N.dumps("WorkUnit") + N.dumps(N.work_unit_tree(WorkUnit(1, 2, (3, 4)))) + b"\x81" + N.dumps("Done"))
    a = run(node, "--decode", str(f1)).stdout.decode().splitlines()
    b = run(node, "--decode", str(f2)).stdout.decode().splitlines()
    assert a[0] == b[0] and a[0].startswith("SetJob id=(1,2) scene=deterministic image=40x30 shapes=5 ")
    assert b[1:] == ["WorkUnit rows=1..2 job=(3,4)", "Done"]
    # half-precision radius
    tree["scene_data"]["shapes"][1]["Sphere"]["radius"] = 1.0
    blob = N.dumps({"SetJob": tree}).replace(N.dumps("radius") + b"\xfa\x3f\x80\x00\x00", N.dumps("radius") + b"\xf9\x3c\x00", 1)
    f1.write_bytes(blob)
    dump = tmp_path / "flat.txt"
    run(node, "--decode-flat", str(f1), str(dump))
    _check_same(_parse_dump(str(dump)), sd)


@pytest.mark.parametrize("mutate,message", [
    (lambda t: t["scene_data"].pop("shapes"), "missing field `shapes`"),
    (lambda t: t["scene_data"]["shapes"][0]["Sphere"].pop("invert"), "missing field `invert`"),
    (lambda t: t["scene_data"]["shapes"].__setitem__(0, {"Torus": {}}), "unknown variant `Torus`"),
    (lambda t: t["config"].__setitem__("sample_root", -1), "expected an unsigned integer"),
    (lambda t: t["scene_data"]["camera_data"].__setitem__("zoom_factor", "wide"), "expected f64"),
    (lambda t: t.__setitem__("id", [1]), "JobID"),
])
def test_malformed_jobs_are_rejected_with_serde_style_messages(node, tmp_path, mutate, message):
    sd = Hp.deterministic_scene(40, 30)
    tree = {"id": [1, 2], "scene_data": N.scene_tree(sd, "map"), "config": {"sample_root": 2, "max_trace_depth": 5, "rows_per_work_unit": 50}}
    mutate(tree)
    f = tmp_path / "bad.cbor"
    f.write_bytes(N.dumps({"SetJob": tree}))
    p = run(node, "--decode", str(f), ok=False)
    assert p.returncode == 101 and message in p.stderr.decode(), p.stderr.decode()


def test_truncated_and_random_input_never_crashes_the_decoder(node, tmp_path):
    sd = Hp.deterministic_scene(40, 30)
    msg = N.set_job((1, 2), sd, JobConfiguration(2, 5, 50))
    f = tmp_path / "x.cbor"
    rng = np.random.default_rng(7)
    cuts = sorted(set(int(c) for c in rng.integers(1, len(msg) - 1, 25)) | {1, 2, len(msg) - 1})
    for c in cuts:
        f.write_bytes(msg[:c])
        p = run(node, "--decode", str(f), ok=False)
        assert p.returncode == 101 and b"end of stream" in p.stderr, (c, p.returncode, p.stderr)
    for k in range(60):   # bit flips and garbage: any outcome but a signal or a hang
        b = bytearray(msg)
        for pos in rng.integers(0, len(b), 1 + k % 4):
            b[pos] = int(rng.integers(0, 256))
        f.write_bytes(bytes(b) if k % 3 else bytes(rng.integers(0, 256, 200, dtype=np.uint8)))
        p = run(node, "--decode", str(f), ok=False)
        assert p.returncode in (0, 101), (k, p.returncode, p.stderr[:200])
    # length fields that promise more than any document may hold must fail fast, not allocate
    for blob in (b"\x9b\xff\xff\xff\xff\xff\xff\xff\xff", b"\xbb\x00\x00\x00\xff\xff\xff\xff\xff", b"\x7b\x7f\xff\xff\xff\xff\xff\xff\xff",
                 b"\x81" * 100_000, b"\xff", b"\x1c"):
        f.write_bytes(blob)
        p = run(node, "--decode", str(f), ok=False)
        assert p.returncode == 101, blob[:10]


def test_rows_ready_encoding_is_exact_for_every_kind_of_double(node):
    """C++ encodes RenderEvent::RowsReady; the Python decoder must get the very same doubles back: values that are
    exact in f32 travel as f32, NaN and the infinities as f16, everything else as f64; -0.0 keeps its sign."""
    rng = np.random.default_rng(11)
    width, n_rows = 9, 4
    px = rng.random(n_rows * width * 3)
    px[:12] = [0.0, -0.0, 1.0, 0.5, float("inf"), float("-inf"), float("nan"), 5e-324, 2.2250738585072014e-308,
               float(np.float32(0.1)), 1.401298464324817e-45, 3.4028234663852886e+38]
    px[12:15] = [3.4028234663852886e+38 * 2, 1e308, 1 + 2 ** -52]
    out = run(node, "--rows-ready", "10", "13", str(width), str(2 ** 64 - 1), "5", stdin=px.tobytes()).stdout
    as_map = run(node, "--enum-form", "map", "--rows-ready", "10", "13", str(width), str(2 ** 64 - 1), "5", stdin=px.tobytes()).stdout
    assert out[:11] == b"\x82\x69RowsReady" and as_map[:11] == b"\xa1\x69RowsReady" and out[1:] == as_map[1:]
    tree, used = N.loads(out)
    assert used == len(out)
    assert tree[0] == "RowsReady" and list(tree[1]) == ["work_unit", "rows"]
    assert list(tree[1]["rows"][0][0]) == ["r", "g", "b"]
    assert N.rows_ready_from_tree(N.loads(as_map)[0]).rows.tobytes() == N.rows_ready_from_tree(tree).rows.tobytes()
    res = N.rows_ready_from_tree(tree)
    assert (res.work_unit.row_start, res.work_unit.row_end, res.work_unit.job_id) == (10, 13, (2 ** 64 - 1, 5))
    got = res.rows.ravel()
    same = (got.view(np.uint64) == px.view(np.uint64)) | (np.isnan(got) & np.isnan(px))
    assert same.all()
    # the first colour on the wire, byte for byte: {"r": 0.0, "g": -0.0, "b": 1.0} as f32
    head = out.index(b"\xa3\x61r")
    assert out[head:head + 22] == bytes.fromhex("a3" "6172" "fa00000000" "6167" "fa80000000" "6162" "fa3f800000")
    # and the Python encoder agrees with the C++ one on the whole message
    again = N.dumps(["RowsReady", {"work_unit": N.work_unit_tree(res.work_unit),
                                   "rows": [[N._color(c) for c in row] for row in px.reshape(n_rows, width, 3)]}])
    assert again == out
    # rows that do not match the width are refused
    p = run(node, "--rows-ready", "0", "0", "7", "0", "0", stdin=px.tobytes(), ok=False)
    assert p.returncode == 101 and b"do not match the width" in p.stderr


def test_node_fails_loudly_without_a_gpu(node):
    """No CPU rendering path: without a CUDA device the server must refuse to start (on a GPU box it would bind; the
    GPU tests cover that)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run(node, "-h", "127.0.0.1", "-p", "0", "--clients", "1", ok=False)
    assert p.returncode == 101 and b"flux_ctx_create" in p.stderr


class FakeNode:
    """The node's end of the protocol without a renderer (test double): answers every WorkUnit with rows whose
    colours are a known function of (row, column), and records what it was sent."""

    def __init__(self, num_threads=3):
        import socket
        import threading
        self.srv = socket.create_server(("127.0.0.1", 0))
        self.port = self.srv.getsockname()[1]
        self.num_threads = num_threads
        self.events, self.error, self.forms = [], None, set()
        self.thread = threading.Thread(target=self._serve, daemon=True)
        self.thread.start()

    @staticmethod
    def colour(row, col):
        return [row / 64.0, col / 32.0, ((row * 31 + col * 17) % 101) / 100.0]

    def _serve(self):
        try:
            conn, _ = self.srv.accept()
            rf = conn.makefile("rb")

            def read(n):
                b = rf.read(n)
                if len(b) != n:
                    raise EOFError
                return b

            conn.sendall(N.worker_info(self.num_threads))
            width = None
            while True:
                req = N.load(read)
                if req == "Done":
                    self.events.append("Done")
                    break
                tag, body = N.variant_parts(req)
                form = "map" if isinstance(req, dict) else "array"
                self.forms.add(form)
                if tag == "SetJob":
                    width = body["scene_data"]["output_settings"]["image_width"]
                    self.events.append(("SetJob", tuple(body["id"]), body["scene_data"]["scene_name"], body["config"]))
                else:
                    self.events.append(("WorkUnit", body["row_start"], body["row_end"], tuple(body["job_id"])))
                    rows = [[dict(zip("rgb", self.colour(r, c))) for c in range(width)] for r in range(body["row_start"], body["row_end"] + 1)]
                    conn.sendall(N.dumps(N.variant("RowsReady", {"work_unit": body, "rows": rows}, form)))
            conn.close()
        except Exception as e:   # surfaced by the test
            self.error = e


@pytest.mark.parametrize("form", ["array", "map"])
def test_cpp_network_worker_drives_a_node(node, tmp_path, form):
    """`fluxb200 -n host:port` (NetworkWorker, workers.rs:118-245): reads WorkerInfo, sends SetJob, the work units of
    Job::work_units in order, Done; assembles the rows it gets back and writes the PPM."""
    import ctypes
    from flux_b200 import _capi
    fake = FakeNode()
    out = tmp_path / "net.ppm"
    p = subprocess.run([os.path.join(ROOT, "host", "fluxb200"), os.path.join(ROOT, "scenes", "demo1.yml"), "-L", "-n", f"127.0.0.1:{fake.port}",
                        "-r", "3", "-d", "4", "-R", "7", "--width", "24", "--height", "20", "--seed", "99", "-o", str(out),
                        "--enum-form", form],
                       capture_output=True, timeout=120)
    fake.thread.join(10)
    assert fake.error is None, fake.error
    assert p.returncode == 0, p.stderr.decode()
    assert b"Threads: 3" in p.stdout
    assert fake.forms == {form}
    assert fake.events[0] == ("SetJob", (99, 0), "demo1", {"sample_root": 3, "max_trace_depth": 4, "rows_per_work_unit": 7})
    assert fake.events[1:-1] == [("WorkUnit", 0, 6, (99, 0)), ("WorkUnit", 7, 13, (99, 0)), ("WorkUnit", 14, 19, (99, 0))]
    assert fake.events[-1] == "Done"
    img = np.array([[FakeNode.colour(r, c) for c in range(24)] for r in range(20)], np.float64)
    ref = tmp_path / "ref.ppm"
    assert _capi.lib().flux_write_ppm(str(ref).encode(), 24, 20, img.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    assert out.read_bytes() == ref.read_bytes()


@pytest.mark.parametrize("form", ["array", "map"])
def test_python_network_worker_drives_a_node(form):
    fake = FakeNode(num_threads=5)
    sd = Hp.deterministic_scene(10, 9)
    w = N.NetworkWorker(f"127.0.0.1:{fake.port}", timeout=30, form=form)
    assert w.info() == {"num_threads": 5}
    img = w.render_job(sd, JobConfiguration(2, 5, 4), job_id=(7, 1))
    fake.thread.join(10)
    assert fake.error is None, fake.error
    assert [e[1:3] for e in fake.events[1:-1]] == [(0, 3), (4, 7), (8, 8)]
    assert np.array_equal(img, np.array([[FakeNode.colour(r, c) for c in range(10)] for r in range(9)]))


def test_cpp_manager_shares_one_job_between_two_nodes(node, tmp_path):
    """`fluxb200 -n a -n b` (flux/src/main.rs:131-137: -n is repeatable): both nodes get the job, the work units are
    handed out one at a time from a single queue (manager.rs:100), every unit is rendered exactly once and the rows
    land where they belong."""
    import ctypes
    from flux_b200 import _capi
    fakes = [FakeNode(num_threads=2), FakeNode(num_threads=7)]
    out = tmp_path / "two.ppm"
    p = subprocess.run([os.path.join(ROOT, "host", "fluxb200"), os.path.join(ROOT, "scenes", "demo1.yml"),
                        "-L", "-n", f"127.0.0.1:{fakes[0].port}", "-n", f"127.0.0.1:{fakes[1].port}",
                        "-r", "2", "-R", "3", "--width", "24", "--height", "31", "--seed", "4", "-o", str(out)],
                       capture_output=True, timeout=120)
    for f in fakes:
        f.thread.join(10)
        assert f.error is None, f.error
    assert p.returncode == 0, p.stderr.decode()
    assert b"Threads: 2" in p.stdout and b"Threads: 7" in p.stdout
    units = []
    for f in fakes:
        assert f.events[0][0] == "SetJob" and f.events[0][1] == (4, 0) and f.events[-1] == "Done"
        units += [e[1:3] for e in f.events[1:-1]]
    assert sorted(units) == [(r, min(30, r + 2)) for r in range(0, 31, 3)]          # each unit once, none lost
    assert all(len(f.events) > 2 for f in fakes)                                       # both nodes took part
    img = np.array([[FakeNode.colour(r, c) for c in range(24)] for r in range(31)], np.float64)
    ref = tmp_path / "ref.ppm"
    assert _capi.lib().flux_write_ppm(str(ref).encode(), 24, 31, img.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    assert out.read_bytes() == ref.read_bytes()


def test_cpp_manager_reports_a_node_that_goes_away(node, tmp_path):
    """A node that closes the connection in the middle of a job is fatal for the job, as in the reference
    (manager.rs:158-161 panics on a lost worker): exit status 101 and a message, no partial file."""
    import socket
    import threading
    srv = socket.create_server(("127.0.0.1", 0))

    def rude():
        conn, _ = srv.accept()
        conn.sendall(N.worker_info(1))
        conn.recv(4096)
        conn.close()

    t = threading.Thread(target=rude, daemon=True)
    t.start()
    out = tmp_path / "gone.ppm"
    p = subprocess.run([os.path.join(ROOT, "host", "fluxb200"), os.path.join(ROOT, "scenes", "demo1.yml"), "-L", "-n", f"127.0.0.1:{srv.getsockname()[1]}",
                        "--width", "8", "--height", "8", "-o", str(out)], capture_output=True, timeout=120)
    t.join(10)
    assert p.returncode == 101 and (b"closed the connection" in p.stderr or b"send:" in p.stderr or b"recv:" in p.stderr), p.stderr
    assert not out.exists()


def test_manager_without_dash_L_needs_a_gpu(node, tmp_path):
    """Without -L the GPUs of the manager's box render beside the nodes (the reference's local worker,
    flux/src/main.rs:43-60); on a box without a GPU that fails loudly instead of falling back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fake = FakeNode()
    p = subprocess.run([os.path.join(ROOT, "host", "fluxb200"), os.path.join(ROOT, "scenes", "demo1.yml"), "-n", f"127.0.0.1:{fake.port}",
                        "--width", "8", "--height", "8", "-o", str(tmp_path / "x.ppm")], capture_output=True, timeout=120)
    assert p.returncode == 101 and b"flux_ctx_create" in p.stderr
