"""world_size-2 `gloo` test of the multi-rank host logic (realization gather, sample sharding)."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

WORKER = textwrap.dedent("""
    import sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch
    from holodeck_b200 import dist
    rank, size = dist.init("gloo")
    assert size == 2 and dist.is_distributed()
    R, F, L = 7, 3, 2
    r0, cnt = dist.realization_slice(R)
    full = torch.arange(F * R * L, dtype=torch.float64).reshape(F, R, L)
    mine = full[:, r0:r0 + cnt].contiguous()
    got = dist.gather_realizations(mine, axis=1, nreals=R)
    assert torch.equal(got, full), (rank, got.shape)
    even = dist.gather_realizations(torch.full((F, 4), float(rank)), axis=1)
    assert even.shape == (F, 8) and float(even[0, 0]) == 0.0 and float(even[0, -1]) == 1.0
    # ONE collective for the two tables of a sam.gwb step
    ss = torch.arange(F * 4 * L, dtype=torch.float64).reshape(F, 4, L) + 1000 * rank
    bg = torch.arange(F * 4, dtype=torch.float64).reshape(F, 4) - 1000 * rank
    (g_ss, g_bg), buf = dist.gather_tables([ss, bg])
    assert g_ss.shape == (F, 8, L) and g_bg.shape == (F, 8)
    for rr in range(2):
        assert torch.equal(g_ss[:, 4*rr:4*rr+4], torch.arange(F * 4 * L, dtype=torch.float64).reshape(F, 4, L) + 1000 * rr)
        assert torch.equal(g_bg[:, 4*rr:4*rr+4], torch.arange(F * 4, dtype=torch.float64).reshape(F, 4) - 1000 * rr)
    (_, _), buf2 = dist.gather_tables([ss, bg], out=buf)
    assert buf2 is buf                                     # the gather buffer is reused from step to step
    # ADVICE r1 (high): with no user seed every rank must still build the SAME space and permutation
    seed = dist.shared_seed(None)
    both_seeds = [None, None]
    torch.distributed.all_gather_object(both_seeds, seed)
    assert both_seeds[0] == both_seeds[1]
    import holodeck_b200 as holo
    space = holo.librarian.PS_Classic_Phenom_Uniform(nsamples=9, sam_shape=8, seed=seed)
    samples = [None, None]
    torch.distributed.all_gather_object(samples, space.param_samples.tolist())
    assert samples[0] == samples[1]
    idx_none = dist.sample_indices(9, seed=seed)
    got = [None, None]
    torch.distributed.all_gather_object(got, [int(v) for v in idx_none])
    assert sorted(got[0] + got[1]) == list(range(9)), got
    assert abs(dist.max_over_ranks(float(rank)) - 1.0) < 1e-12
    idx = dist.sample_indices(11, seed=5)
    both = [torch.zeros(6, dtype=torch.int64) for _ in range(2)]
    pad = torch.full((6,), -1, dtype=torch.int64); pad[:len(idx)] = torch.as_tensor(idx)
    torch.distributed.all_gather(both, pad)
    allidx = sorted(int(v) for t in both for v in t if v >= 0)
    assert allidx == list(range(11)), allidx
    dist.barrier()
    dist.finalize()
    assert not dist.is_distributed()
    dist.finalize()          # idempotent; a no-op without a process group
    print("ok", rank)
""")


def test_two_rank_gather_and_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-3000:]
    assert res.stdout.count("ok") == 2
