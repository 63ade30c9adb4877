"""N>1 host logic on the CPU: two gloo ranks shard a list of sets, each runs its slice (through the
kernel single-stepper tests/emu -- logic only), rank 0 gathers; the result equals the oracle's."""
import os
import subprocess
import sys

from common import ROOT
from csa_b200.shard import shard_bounds

WORKER = r'''
import os, sys, random, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
from common import EMU_LIB, build_emu, gen_case, oracle_run
from csa_b200.api import RotationFinder
from csa_b200.shard import find_rotations_sharded
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = random.Random(123)
sets = [gen_case(rng, max_n=400)[1] for _ in range(11)]
rf = RotationFinder(lib_path=EMU_LIB)
out = find_rotations_sharded(rf, sets, rank, world, dist)
if rank == 0:
    assert len(out) == len(sets)
    for (status, rot), s in zip(out, sets):
        o = oracle_run(s)
        assert status == o["status"]
        if status == 0:
            assert rot == list(o["rotations"])
    print("SHARD_OK", world)
dist.destroy_process_group()
'''


def test_shard_bounds():
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(3, 8) == [0, 1, 2, 3, 3, 3, 3, 3, 3]
    assert shard_bounds(0, 2) == [0, 0, 0]


def test_two_gloo_ranks(tmp_path):
    from common import build_emu
    build_emu()
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(w), ROOT],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=env, text=True)
    assert p.returncode == 0 and "SHARD_OK 2" in p.stdout, p.stdout[-2000:]


def test_stream_in_batches(emu_finder):
    import random
    from common import compare_with_oracle, gen_case, oracle_run
    from csa_b200.shard import batches_by_size, find_rotations_stream
    rng = random.Random(9)
    sets = [gen_case(rng, max_n=450)[1] for _ in range(23)]
    cuts = list(batches_by_size(sets, 3000))
    assert len(cuts) > 3 and sum(len(c) for _, c in cuts) == len(sets) and [s for s, _ in cuts] == sorted(s for s, _ in cuts)
    res = find_rotations_stream(emu_finder, sets, max_bases=3000, with_blocks=True)
    for i, (r, s) in enumerate(zip(res, sets)):
        compare_with_oracle(r, oracle_run(s), s, f"stream set {i}")
