"""fluxb200-node on a real GPU: a manager speaking the reference's protocol (workers.rs:118-245) gets back exactly
the rows the GPU worker renders locally.  Both clients are exercised: the Python NetworkWorker (independent codec)
and the C++ one (`fluxb200 -n`)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from flux_b200 import JobConfiguration, SceneData, _capi
from flux_b200 import netproto as N
from flux_b200.worker import GpuWorker
from tests import helpers as Hp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def start_node(*extra):
    subprocess.run(["make", "-C", os.path.join(ROOT, "host"), "-s"], check=True)
    p = subprocess.Popen([os.path.join(ROOT, "host", "fluxb200-node"), "-h", "127.0.0.1", "-p", "0", "--clients", "1", *extra],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    line = p.stdout.readline().decode()
    assert line.startswith("Bind address: 127.0.0.1:"), (line, p.stderr.read().decode() if p.poll() is not None else "")
    return p, int(line.rsplit(":", 1)[1])


@pytest.mark.parametrize("form", ["array", "map"])
def test_python_manager_gets_the_rows_the_gpu_renders(gpu_ctx, form):
    """The node answers in the enum form its manager writes (serde_cbor < 0.10 reads only the array form)."""
    sd = Hp.deterministic_scene(96, 64)
    cfg = JobConfiguration(4, 5, 10)
    proc, port = start_node("--seed", "5")
    try:
        w = N.NetworkWorker(f"127.0.0.1:{port}", timeout=120, form=form)
        assert w.info() == {"num_threads": 1}
        seen = []
        raw_load = N.load
        def spy(read, _depth=0):   # what arrives, before NetworkWorker interprets it
            v = raw_load(read, _depth)
            if _depth == 0:
                seen.append(type(v))
            return v
        N.load = spy
        try:
            img = w.render_job(sd, cfg, job_id=(2 ** 64 - 1, 9))
        finally:
            N.load = raw_load
        assert set(seen) == {list if form == "array" else dict}
        assert proc.wait(30) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
    log = proc.stdout.read().decode()
    assert "Got connection from 127.0.0.1:" in log and "Got job" in log and "Got done message, shutting down" in log
    local = GpuWorker(0, seed=5)
    try:
        ref = local.render_image(sd, cfg)
    finally:
        local.stop()
    assert np.array_equal(img.view(np.uint64), ref.view(np.uint64))
    assert float(ref.max()) > 0.1


def test_cpp_manager_over_the_node_writes_the_local_ppm(tmp_path):
    """`fluxb200 demo2.yml -n node` == `fluxb200 demo2.yml` byte for byte (same seed, glossy scene, depth of field)."""
    cli = os.path.join(ROOT, "host", "fluxb200")
    scene = os.path.join(ROOT, "scenes", "demo2.yml")
    common = ["-r", "4", "-R", "9", "--width", "64", "--height", "48", "--seed", "3"]
    proc, port = start_node("--seed", "3")
    try:
        a = subprocess.run([cli, scene, "-L", "-n", f"127.0.0.1:{port}", *common, "-o", str(tmp_path / "net.ppm")], capture_output=True, timeout=300)
        assert a.returncode == 0, a.stderr.decode()
        assert proc.wait(30) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
    b = subprocess.run([cli, scene, *common, "-o", str(tmp_path / "local.ppm")], capture_output=True, timeout=300)
    assert b.returncode == 0, b.stderr.decode()
    assert (tmp_path / "net.ppm").read_bytes() == (tmp_path / "local.ppm").read_bytes()


def test_node_drops_a_client_that_breaks_the_protocol_and_serves_the_next():
    import socket
    proc, port = start_node_n(2)
    try:
        s = socket.create_connection(("127.0.0.1", port), timeout=30)
        rf = s.makefile("rb")
        assert N.load(lambda n: rf.read(n)) == {"num_threads": 1}
        s.sendall(N.work_unit(N.WorkUnit(0, 3, (1, 1))))       # a unit before any job
        assert rf.read(1) == b""                                # dropped
        s.close()
        sd = Hp.deterministic_scene(32, 16)
        w = N.NetworkWorker(f"127.0.0.1:{port}", timeout=120)
        img = w.render_job(sd, JobConfiguration(2, 5, 16))
        assert img.shape == (16, 32, 3) and float(img.max()) > 0.1
        assert proc.wait(30) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
    assert "handle_client exited with work unit before SetJob" in proc.stdout.read().decode()


def start_node_n(n):
    subprocess.run(["make", "-C", os.path.join(ROOT, "host"), "-s"], check=True)
    p = subprocess.Popen([os.path.join(ROOT, "host", "fluxb200-node"), "-h", "127.0.0.1", "-p", "0", "--clients", str(n)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    line = p.stdout.readline().decode()
    assert line.startswith("Bind address: 127.0.0.1:"), line
    return p, int(line.rsplit(":", 1)[1])


def test_sharded_worker_paths_render_the_same_pixels(tmp_path):
    """Two contexts (here on the same GPU, `--devices 0,0`) exercise the multi-GPU paths of the C++ host: work units
    sharded in interleaved tiles by the node (GpuWorker::render_unit) and the whole frame sharded by the driver
    (GpuWorker::render_job).  Pixels must not depend on the sharding."""
    cli = os.path.join(ROOT, "host", "fluxb200")
    scene = os.path.join(ROOT, "scenes", "demo2.yml")
    common = ["-r", "4", "-R", "11", "--width", "64", "--height", "50", "--seed", "8"]
    one = subprocess.run([cli, scene, *common, "-o", str(tmp_path / "one.ppm")], capture_output=True, timeout=300)
    assert one.returncode == 0, one.stderr.decode()
    two = subprocess.run([cli, scene, *common, "--devices", "0,0,0", "-o", str(tmp_path / "three.ppm")], capture_output=True, timeout=300)
    assert two.returncode == 0, two.stderr.decode()
    assert b"on 3 GPU(s)" in two.stdout
    assert (tmp_path / "three.ppm").read_bytes() == (tmp_path / "one.ppm").read_bytes()
    proc, port = start_node("--seed", "8", "--devices", "0,0")
    try:
        net = subprocess.run([cli, scene, "-L", "-n", f"127.0.0.1:{port}", *common, "-o", str(tmp_path / "net.ppm")], capture_output=True, timeout=300)
        assert net.returncode == 0, net.stderr.decode()
        assert b"Threads: 2" in net.stdout
        assert proc.wait(30) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
    assert (tmp_path / "net.ppm").read_bytes() == (tmp_path / "one.ppm").read_bytes()


def test_local_gpus_and_a_node_share_one_job(tmp_path):
    """Without -L the manager's own GPUs pull work units from the same queue as the node (LocalWorker beside
    NetworkWorker in the reference).  Same seed everywhere, so the frame is the local render's, whoever rendered
    which unit; both must have rendered some."""
    cli = os.path.join(ROOT, "host", "fluxb200")
    scene = os.path.join(ROOT, "scenes", "demo2.yml")
    common = ["-r", "16", "-R", "5", "--width", "96", "--height", "80", "--seed", "6"]
    proc, port = start_node("--seed", "6")
    try:
        both = subprocess.run([cli, scene, "-n", f"127.0.0.1:{port}", *common, "-o", str(tmp_path / "both.ppm")], capture_output=True, timeout=300)
        assert both.returncode == 0, both.stderr.decode()
        assert proc.wait(30) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
    assert both.stdout.count(b"ready, info:") == 2
    local = subprocess.run([cli, scene, *common, "-o", str(tmp_path / "local.ppm")], capture_output=True, timeout=300)
    assert local.returncode == 0, local.stderr.decode()
    assert (tmp_path / "both.ppm").read_bytes() == (tmp_path / "local.ppm").read_bytes()
