"""Multi-rank host logic on CPU: the slab-decomposition plan (mb_decomp_plan, the code the decomposed step uses to
post its NCCL send/recv) driven through torch.distributed with the gloo backend, world_size 2 (and 3).

Every rank owns a slab of cell layers; after the planned exchange each rank must hold exactly the positions of
the h = 2 layers below and above its slab, the pairs (sender, receiver) must post matching segment sequences, and
the all-gather used at rebuilds must reproduce the full array."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mollyb200 as mb


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _layer_lo(q, ncz, p):
    return (q * ncz) // p


def _worker(rank, world, port, ncz, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)  # same stream on every rank: the replicated sort is identical
    counts = rng.integers(0, 40, ncz)
    layer_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n = int(layer_start[-1])
    truth = rng.normal(size=(n, 4)).astype(np.float32)  # the "pos4" array after K1 on the owners
    lo, hi = _layer_lo(rank, ncz, world), _layer_lo(rank + 1, ncz, world)
    mine = torch.full((n, 4), float("nan"))
    mine[layer_start[lo]:layer_start[hi]] = torch.from_numpy(truth[layer_start[lo]:layer_start[hi]])
    send, recv = mb.decomp_plan(ncz, 2, world, rank, layer_start)
    reqs = []
    for peer, st, cnt in send:
        if cnt > 0:
            reqs.append(dist.isend(mine[st:st + cnt].contiguous(), peer))
    bufs = []
    for peer, st, cnt in recv:
        if cnt > 0:
            b = torch.empty((cnt, 4))
            bufs.append((st, cnt, b))
            reqs.append(dist.irecv(b, peer))
    for r in reqs:
        r.wait()
    for st, cnt, b in bufs:
        mine[st:st + cnt] = b
    # needed layers are present and correct, everything else is still unknown
    need = set()
    for l in list(range(lo - 2, lo)) + list(range(hi, hi + 2)):
        need.add(l % ncz)
    need -= set(range(lo, hi))
    ok = True
    for l in range(ncz):
        seg = mine[layer_start[l]:layer_start[l + 1]].numpy()
        ref = truth[layer_start[l]:layer_start[l + 1]]
        if l in need or lo <= l < hi:
            ok &= bool(np.array_equal(seg, ref))
        else:
            ok &= bool(np.isnan(seg).all())
    # rebuild-time replication: broadcast every owner's segment
    full = mine.clone()
    for q in range(world):
        a, b = layer_start[_layer_lo(q, ncz, world)], layer_start[_layer_lo(q + 1, ncz, world)]
        if b > a:
            t = full[a:b].contiguous()
            dist.broadcast(t, src=q)
            full[a:b] = t
    ok &= bool(np.array_equal(full.numpy(), truth))
    # global momentum-style reduction is identical on every rank
    part = torch.tensor([float(truth[layer_start[lo]:layer_start[hi]].astype(np.float64).sum())], dtype=torch.float64)
    dist.all_reduce(part)
    ok &= abs(part.item() - truth.astype(np.float64).sum()) < 1e-6
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "fail")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ncz", [(2, 35), (2, 5), (3, 7)])
def test_slab_plan_exchange_gloo(tmp_path, world, ncz):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ncz, 1234 + ncz, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"


def test_plan_partition_properties():
    for world in (1, 2, 4, 8):
        for ncz in (8, 35, 55):
            ls = np.arange(ncz + 1, dtype=np.int32) * 7
            owned = []
            for r in range(world):
                lo, hi = _layer_lo(r, ncz, world), _layer_lo(r + 1, ncz, world)
                owned += list(range(lo, hi))
                send, recv = mb.decomp_plan(ncz, 2, world, r, ls)
                assert all(p != r for p, _, _ in send + recv)
                if world == 1:
                    assert send == [] and recv == []
            assert owned == list(range(ncz))
            # what r receives from q is what q sends to r, in the same order
            plans = [mb.decomp_plan(ncz, 2, world, r, ls) for r in range(world)]
            for r in range(world):
                for q in range(world):
                    got = [(s, c) for p, s, c in plans[r][1] if p == q]
                    sent = [(s, c) for p, s, c in plans[q][0] if p == r]
                    assert got == sent
