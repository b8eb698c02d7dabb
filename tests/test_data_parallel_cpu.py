"""CPU, gloo, world_size 2: the host-side data-parallel logic (shard bounds, flat gradient
buffer, one collective per step, rank-averaged gradients == global-batch gradient)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from objectpermanence_b200.data_parallel import FlatGradAllReducer, broadcast_parameters, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # deliberately different initial weights per rank
    model = torch.nn.Sequential(torch.nn.Linear(6, 5, bias=False), torch.nn.Tanh(), torch.nn.Linear(5, 4))
    broadcast_parameters(model.parameters(), src=0)
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 4, generator=g)
    lo, hi = shard_bounds(8, rank, world)
    reducer = FlatGradAllReducer(model.parameters())
    loss = (model(x[lo:hi]) - y[lo:hi]).abs().mean()
    loss.backward()
    flat = reducer.reduce()
    assert reducer.collectives == 1
    # every p.grad is a view into the one flat buffer
    assert all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in model.parameters())
    torch.save({"flat": flat.clone(), "params": [p.detach().clone() for p in model.parameters()]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["flat"], r1["flat"])                      # identical on every rank
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)                                    # broadcast worked
    # single-process gradient of the global batch
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5, bias=False), torch.nn.Tanh(), torch.nn.Linear(5, 4))
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 4, generator=g)
    (model(x) - y).abs().mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(r0["flat"], want, atol=1e-6)


def test_shard_bounds():
    assert [shard_bounds(256, r, 8) for r in (0, 7)] == [(0, 32), (224, 256)]
    with pytest.raises(ValueError):
        shard_bounds(30, 0, 4)


def test_reducer_single_process_is_identity():
    lin = torch.nn.Linear(3, 2)
    lin(torch.ones(4, 3)).sum().backward()
    want = torch.cat([p.grad.reshape(-1).clone() for p in lin.parameters()])
    r = FlatGradAllReducer(lin.parameters())
    flat = r.reduce()
    assert torch.equal(flat, want) and r.collectives == 0 and r.nbytes == 4 * 8
