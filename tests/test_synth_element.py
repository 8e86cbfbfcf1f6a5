"""The device generator's element function (csrc/synth_element.h: 32-bit pair indices, one wide multiply, split
53-bit conversion) is the numpy generator (synth.counter_eri_element) bit for bit: host build, CPU only."""
import os
import struct
import subprocess

import numpy as np

import jues.jl_b200 as jb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_build_matches_numpy_generator(tmp_path):
    exe = str(tmp_path / "synth_element_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "synth_element_test.cpp")],
                   check=True)
    rng = np.random.default_rng(11)
    quads = []
    for n in (7, 120, 500, 4000, 65535):
        q = rng.integers(0, n, size=(4000, 4))
        q[:50, 0] = q[:50, 1]                      # diagonal pairs
        q[50:100, 2:] = n - 1                      # the largest pair index
        quads.append(q)
    q = np.concatenate(quads).astype(np.uint64)
    for seed, scale in ((2024, 0.0123), (0, 1.0), (2**63 + 12345, 3.5e-4)):
        inp = f"{seed} {scale!r}\n" + "\n".join(" ".join(str(int(x)) for x in row) for row in q) + "\n"
        r = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True)
        got = np.array([struct.unpack("<d", bytes.fromhex(h)[::-1])[0] for h in r.stdout.split()])
        ref = jb.synth.counter_eri_element(q[:, 0], q[:, 1], q[:, 2], q[:, 3], seed, scale)
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint64), np.asarray(ref, dtype=np.float64).view(np.uint64))
