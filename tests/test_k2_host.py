"""CPU: the K2 arithmetic of the CUDA engine (envelope runs, leaf plans, add-chain memo, single-boundary
and row-wise mixed leaves, tree walk) compiled for the HOST from the very header the kernels include, against
a literal per-site numpy-pairwise evaluation (waafle/waafle_orgscorer.py:371-382,403).  Bit-exact.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
HDR = os.path.join(os.path.dirname(HERE), "waafle_b200", "csrc", "wfl_warp_common.cuh")


def test_k2_region_on_host(tmp_path):
    src = open(HDR).read()
    region = src[src.index("// K2-HOST-BEGIN"):src.index("// K2-HOST-END")]
    (tmp_path / "k2_region.inc").write_text(region)
    exe = str(tmp_path / "k2_host")
    cc = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cc, "-O1", "-ffp-contract=off", "-std=c++17", "-I", str(tmp_path),
                    os.path.join(HERE, "host_k2", "k2_host.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "30000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout
