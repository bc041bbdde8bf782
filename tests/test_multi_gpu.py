"""Multi-GPU path: world_size-2 parity against the oracle on a box with >= 2 GPUs (tests/multi_gpu_parity.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_parity():
    import hga_b200
    n = hga_b200.capi.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_parity.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("multi-GPU parity ok") == 2
    assert r.stdout.count("gathered handle ok") == 2
    assert r.stdout.count("fault injection ok") == 1
    assert r.stdout.count("engine mirror ok") == 1


@pytest.mark.gpu
def test_cli_two_gpus_exports_the_single_gpu_components(tmp_path):
    """categorization --gpus 2 (rank threads + NCCL + gather to the first GPU) writes the files categorization writes on one GPU"""
    import hga_b200
    import datagen
    if hga_b200.capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "hybrid-genome-assembler_b200", "categorization")
    a = datagen.random_genome(60000, 901)
    b = datagen.mutate(a, 0.02, 902)
    ra = datagen.sample_reads(a, 400, 1500, 31, error_rate=0.03, length_sigma=0.5, max_len=60000)
    rb = datagen.sample_reads(b, 400, 1500, 32, error_rate=0.03, length_sigma=0.5, max_len=60000)
    fa, fb, kp = str(tmp_path / "a.fa"), str(tmp_path / "b.fa"), str(tmp_path / "19-mers.txt")
    datagen.write_fasta(fa, ra, prefix="a_"); datagen.write_fasta(fb, rb, prefix="b_")
    datagen.write_kmers(kp, datagen.discriminative_kmers([a, b], 19), 19)
    outs = []
    for gpus in (1, 2):
        out = str(tmp_path / f"out{gpus}")
        r = subprocess.run([exe, "--kmers", kp, "-o", out, "--sc_min_size", "5", "--core_enrichment", "5", "--tail_amplification", "10", "--gpus", str(gpus), fa, fb],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        for label in ("Index construction took", "Calculation of connections between reads took", "Union-find took", "Exported"):
            assert label in r.stdout, (label, r.stdout)
        outs.append({f: open(os.path.join(out, f), "rb").read() for f in sorted(os.listdir(out))})
    assert outs[0].keys() == outs[1].keys() and len(outs[0]) > 0
    for f in outs[0]:
        assert outs[0][f] == outs[1][f], f"component file {f} differs between 1 and 2 GPUs"
