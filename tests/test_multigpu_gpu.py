"""N>1 on real GPUs (NCCL): sharded drivers equal the single-GPU results bit for bit.
Skipped on single-GPU boxes; the host logic is covered on CPU by tests/test_sharding.py (gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import cloud

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import geometric_adv_b200 as ga
        from geometric_adv_b200 import sharding
        from geometric_adv_b200.attack import PointNetAE, attack_pairs
        dev = torch.device("cuda", rank)
        clouds = torch.from_numpy(cloud(3, (37, 512, 3))).to(dev)
        cd = sharding.all_pairs_chamfer(clouds)
        sl = [0, 5, 20, 37]
        cd2, nn2 = sharding.prepare_indices(clouds, sl)
        clouds_even = torch.from_numpy(cloud(7, (16, 300, 3))).to(dev)  # equal blocks: all_gather_into_tensor path
        cd3, nn3 = sharding.prepare_indices(clouds_even, [0, 9, 16])
        pc = torch.from_numpy(cloud(4, (21, 1024, 3))).to(dev)
        kd = sharding.knn_dists_sharded(pc, 10)
        torch.manual_seed(0)
        ae = PointNetAE(256)
        src, tgt = torch.from_numpy(cloud(5, (8, 256, 3))), torch.from_numpy(cloud(6, (8, 256, 3)))
        m, a = attack_pairs(ae, src, tgt, batch_size=2, num_iterations=5, num_iterations_thresh=2, device=str(dev))
        if rank == 0:
            full = ga.chamfer_all_pairs(clouds)
            kd1 = ga.knn_dists(pc, 10)
            ret["cd_equal"] = bool(torch.equal(cd, full))
            ret["prep_equal"] = bool(torch.equal(cd2, full)) and np.array_equal(
                nn2.cpu().numpy(), sharding.sort_dist_mat(full, sl))
            full3 = ga.chamfer_all_pairs(clouds_even)
            ret["prep_even_equal"] = bool(torch.equal(cd3, full3)) and np.array_equal(
                nn3.cpu().numpy(), sharding.sort_dist_mat(full3, [0, 9, 16]))
            ret["kd_equal"] = bool(torch.equal(kd, kd1))
            ret["attack"] = (m.cpu().numpy(), a.cpu().numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_two_rank_nccl_equals_single_gpu():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["cd_equal"], "sharded all-pairs matrix differs from the single-GPU one"
    assert ret["kd_equal"], "sharded kNN distances differ from the single-GPU ones"
    assert ret["prep_equal"] and ret["prep_even_equal"], "sharded prepare_indices differs from the single-GPU files"
    from geometric_adv_b200.attack import PointNetAE, attack_pairs
    torch.manual_seed(0)
    ae = PointNetAE(256)
    src, tgt = torch.from_numpy(cloud(5, (8, 256, 3))), torch.from_numpy(cloud(6, (8, 256, 3)))
    m1, a1 = attack_pairs(ae, src, tgt, batch_size=2, num_iterations=5, num_iterations_thresh=2)
    m2, a2 = ret["attack"]
    assert np.array_equal(m1.cpu().numpy(), m2) and np.array_equal(a1.cpu().numpy(), a2)
