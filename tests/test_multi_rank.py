"""N > 1 path on CPU: two ranks (gloo), each with its own engine (the emulator build of the
kernel sources) and its shard of the call legs; the gathered outputs must equal a single
engine that serves all legs -- legs shard with no data-path collective."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
N_LEGS, N_FRAMES = 4, 40


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_path, ret):
    for p in (os.path.join(ROOT, "webrtc-audio-processing_b200", "python"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import wap_b200
    import wap_shard
    from common import run_legs, synthetic_leg
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lib = wap_b200.load(emu_path)
    mine = wap_shard.legs_of_rank(N_LEGS, rank, world)
    legs = [synthetic_leg(i, N_FRAMES) for i in mine]
    out, _ = run_legs(lib, 16000, legs, aec=True, ns=True, ns_level=1)
    gathered = [None] * world
    dist.all_gather_object(gathered, [o for o in out])   # result hand-back only: not on the data path
    dist.barrier()
    if rank == 0:
        ret["merged"] = wap_shard.merge_outputs(gathered, world)
    dist.destroy_process_group()


def test_two_ranks_shard_legs_without_data_path_collective(emu_lib):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from common import run_legs, synthetic_leg
    emu_path = build_emu.build(verbose=False)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, emu_path, ret), nprocs=2, join=True)
    merged = ret["merged"]
    single, _ = run_legs(emu_lib, 16000, [synthetic_leg(i, N_FRAMES) for i in range(N_LEGS)],
                         aec=True, ns=True, ns_level=1)
    assert len(merged) == N_LEGS
    for i in range(N_LEGS):
        assert np.array_equal(np.asarray(merged[i]), single[i]), i


def test_shard_helpers():
    sys.path.insert(0, os.path.join(ROOT, "webrtc-audio-processing_b200", "python"))
    import wap_shard
    for world in (1, 2, 4, 8):
        seen = sorted(l for r in range(world) for l in wap_shard.legs_of_rank(37, r, world))
        assert seen == list(range(37))
        assert all(wap_shard.rank_of_leg(l, world) == r for r in range(world) for l in wap_shard.legs_of_rank(37, r, world))
