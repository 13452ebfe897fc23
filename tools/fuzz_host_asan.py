#!/usr/bin/env python
"""Fuzz the host-side readers under AddressSanitizer + UBSan (no GPU): mutated YAML scenes through
`fluxb200 --dump-flat`, mutated CBOR requests (both enum forms) through `fluxb200-node --decode / --reencode`.
Any exit status other than 0 (accepted) or 101 (rejected with a message), or any sanitizer report, is a failure.
usage: tools/fuzz_host_asan.py [iterations]   (round 1: 500 YAML + 800 CBOR runs, no findings)"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from flux_b200 import JobConfiguration, WorkUnit, netproto as N  # noqa: E402
from tests import helpers as Hp  # noqa: E402
from tests.test_cpp_host import EXT_YAML  # noqa: E402


def build(tmp):
    src = ["fluxhost.cpp", "fluxnet.cpp", "cbor.cpp"]
    flags = ["-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-pthread"]
    link = ["-L" + os.path.join(ROOT, "flux_b200", "lib"), "-lfluxb200", "-Wl,-rpath," + os.path.join(ROOT, "flux_b200", "lib")]
    out = {}
    for name, main in (("cli", "main.cpp"), ("node", "node_main.cpp")):
        out[name] = os.path.join(tmp, name)
        subprocess.run(["g++", *flags, "-o", out[name], main, *src, *link], cwd=os.path.join(ROOT, "host"), check=True)
    return out


def run(cmd, env):
    p = subprocess.run(cmd, capture_output=True, env=env, timeout=120)
    ok = p.returncode in (0, 101) and b"Sanitizer" not in p.stderr and b"runtime error" not in p.stderr
    if not ok:
        print("FAIL", cmd, p.returncode, p.stderr[-800:].decode(errors="replace"))
    return ok


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    rng = np.random.default_rng(5)
    env = dict(os.environ, UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        bins = build(tmp)
        yamls = [open(os.path.join(ROOT, "scenes", "demo2.yml"), "rb").read(), EXT_YAML.encode()]
        f = os.path.join(tmp, "fz.yml")
        for k in range(iters):
            b = bytearray(yamls[k % 2])
            m = k % 5
            if m == 0:
                for pos in rng.integers(0, len(b), 1 + k % 6):
                    b[pos] = int(rng.integers(32, 127))
            elif m == 1:
                b = b[:int(rng.integers(1, len(b)))]
            elif m == 2:
                lines = bytes(b).split(b"\n"); del lines[int(rng.integers(0, len(lines)))]; b = bytearray(b"\n".join(lines))
            elif m == 3:
                lines = bytes(b).split(b"\n"); i = int(rng.integers(0, len(lines)))
                lines[i] = b" " * int(rng.integers(0, 9)) + lines[i].lstrip(); b = bytearray(b"\n".join(lines))
            else:
                for pos in rng.integers(0, len(b), 3):
                    b[pos] = int(rng.choice(list(b"[]{}:,&*-#'\"\n ")))
            open(f, "wb").write(bytes(b))
            bad += not run([bins["cli"], f, "--dump-flat", os.path.join(tmp, "flat.txt")], env)
        sd = Hp.deterministic_scene(40, 30)
        msgs = [N.set_job((1, 2), sd, JobConfiguration(2, 5, 50), form) + N.work_unit(WorkUnit(0, 3, (1, 2)), form) + N.done()
                for form in ("array", "map")]
        f = os.path.join(tmp, "fz.cbor")
        for k in range(iters):
            b = bytearray(msgs[k % 2])
            m = k % 4
            if m == 0:
                for pos in rng.integers(0, len(b), 1 + k % 5):
                    b[pos] = int(rng.integers(0, 256))
            elif m == 1:
                b = b[:int(rng.integers(1, len(b)))]
            elif m == 2:
                pos = int(rng.integers(0, len(b))); b[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8))
            else:
                b = bytearray(rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8))
            open(f, "wb").write(bytes(b))
            for tool in ("--decode", "--reencode"):
                bad += not run([bins["node"], tool, f], env)
    print(f"{iters} YAML + {2 * iters} CBOR runs, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
