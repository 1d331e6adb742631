"""CPU test of the block FFT core (genetic_b200/csrc/fft_core.cuh compiled for the host): every per-thread stage function
is executed for every thread in sequence, stage by stage -- what the __syncthreads-separated device code does -- and
compared with a long-double naive DFT. Covers BlockFFT (all sizes, both directions, both precisions) and DitRows, the
natural-order row transform behind the c2r pass. No GPU needed."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_block_fft_and_dit_rows_on_the_host(tmp_path):
    src = os.path.join(HERE, "host", "fft_core_host_test.cu")
    exe = str(tmp_path / "fft_core_host_test")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-ccbin", "/usr/bin/g++", "-gencode",
                           "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "-diag-suppress", "128,20013",
                           "-o", exe, src], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "ALL OK" in out.stdout and out.stdout.count("DitRows") == 4


def test_mixed_radix_stockham_stage_on_the_host(tmp_path):
    """csrc/fft_generic_core.cuh (the line transform for grid sizes that are not powers of two) compiled for the host:
    17 sizes with radices 4, 2, 3, 5, 7 and primes up to 17, both directions, against a long-double DFT."""
    src = os.path.join(HERE, "host", "fft_generic_host_test.cu")
    exe = str(tmp_path / "fft_generic_host_test")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-ccbin", "/usr/bin/g++", "-gencode",
                           "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "-diag-suppress", "128,20013",
                           "-o", exe, src], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "ALL OK" in out.stdout and out.stdout.count("generic n=") == 34
