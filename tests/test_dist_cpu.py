"""CPU, world_size 2 over gloo: the N>1 path's only collective (all-gather of box records)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from feartracker_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.shard_range(total, rank, world)
    full = torch.arange(total * 48, dtype=torch.int64).remainder(251).to(torch.uint8).reshape(total, 48)
    gathered = sharding.all_gather_boxes(full[b:e].clone(), total)
    ok = gathered.shape == full.shape and torch.equal(gathered, full)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok), b, e]))
    dist.barrier()
    dist.destroy_process_group()


def _run(total, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    res = [np.load(os.path.join(tmp_path, f"ok{r}.npy")) for r in range(2)]
    assert all(r[0] == 1 for r in res), res
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total


def test_all_gather_boxes_even(tmp_path):
    _run(8, tmp_path)


def test_all_gather_boxes_ragged(tmp_path):
    _run(7, tmp_path)
