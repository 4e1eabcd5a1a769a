"""The host-side copies the result scatter uses (bamsignals_b200/csrc/hostsimd.cpp: AVX2 widening of the byte-packed
result to int32 and int32 copies, both with streaming stores) against plain loops, for every small size and alignment,
under ASan/UBSan.  CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_widen_and_copy(tmp_path):
    exe = str(tmp_path / "simd")
    subprocess.check_call(["g++", "-O2", "-g", "-fsanitize=address,undefined", "-o", exe,
                           os.path.join(ROOT, "tests", "host_simd_harness.cpp"),
                           os.path.join(ROOT, "bamsignals_b200", "csrc", "hostsimd.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok ")
