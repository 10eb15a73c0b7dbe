"""Host-side checks of the tensor-memory screening kernel's shared-memory layout and stage-ring rule (no GPU needed: the
functions are __host__ __device__ in tau_group_tc_kernel.cuh and nvcc builds a host program from them)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_tc_layout_fits_and_ring_never_overwrites(tmp_path):
    exe = tmp_path / "probe"
    r = subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe),
                        os.path.join(ROOT, "tests", "tc_layout_probe.cu")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout[-2000:]
    lines = out.stdout.split("\n")
    assert "ring ok" in lines
    n = 0
    for ln in lines:
        f = ln.split()
        if not f or f[0] != "layout":
            continue
        S, G, SK, nkb, NC, total, nst, ntb, nacc, stage_b, table_b, tmem = map(int, f[1:])
        n += 1
        assert SK % 4 == 0 and SK * nkb >= S and SK <= 64 + 3
        assert NC % 8 == 0 and NC >= 3 * G
        assert total + 2048 <= 227 * 1024                      # dynamic shared memory of one CTA (+ the static part)
        assert nst >= 2 and ntb >= 2 and nacc >= 2             # double buffering everywhere at least
        assert stage_b == 16 * (SK // 2) * 128                 # 128 rows in granules of 8 rows x (SK/2 chunks) x 16 bytes
        assert table_b == (2 * NC // 8) * (SK // 2) * 128
        assert 32 <= tmem <= 512 and tmem & (tmem - 1) == 0    # tensor-memory allocation: a power of two of columns
    assert n > 1000
    # the BASELINE shapes fit: C3 (S=64, G=8), C4 (S=256, G=16), C5 (S=128, G=20)
    assert [ln for ln in lines if ln.startswith("baseline")] == ["baseline 64 8 1", "baseline 256 16 1", "baseline 128 20 1"]
