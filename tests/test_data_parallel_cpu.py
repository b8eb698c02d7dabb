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


class _FakeWork:
    def __init__(self, log):
        self.log = log

    def wait(self):
        self.log.append(("wait",))


def _stub_nccl(monkeypatch, log):
    """dist.* as a 2-rank NCCL job would answer, with all_reduce recording the slice it was handed."""
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_world_size", lambda group=None: 2)
    monkeypatch.setattr(dist, "get_backend", lambda group=None: "nccl")

    def all_reduce(t, op=None, group=None, async_op=False):
        log.append(("all_reduce", t.data_ptr(), t.numel(), bool(async_op), op))
        return _FakeWork(log) if async_op else None

    monkeypatch.setattr(dist, "all_reduce", all_reduce)


def test_early_bucket_covers_exactly_the_final_gradients_and_reduce_covers_the_rest(monkeypatch):
    """Host logic of the two-bucket all-reduce (OPNet: the weight gradients of LSTM2 are final inside the backward pass): the hook
    starts ONE asynchronous all-reduce over the contiguous slice the gradients occupy, reduce() sends the two ends of the buffer
    and then waits for the bucket; gradients that are not back-to-back slices of the buffer are left to reduce()."""
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 7, 11, 3)]
    red = FlatGradAllReducer(params)
    for p, v in zip(params, red._views):
        v.fill_(1.0)
        p.grad = v
    log = []
    _stub_nccl(monkeypatch, log)
    base, esz = red.flat.data_ptr(), red.flat.element_size()

    # not contiguous in the buffer (views 0 and 2): nothing starts
    red._start_early_bucket((red._views[0], red._views[2]))
    assert red._early is None and log == []
    # a tensor that is not part of the buffer: nothing starts
    red._start_early_bucket((torch.zeros(7),))
    assert red._early is None and log == []

    # views 1 and 2, handed over in any order: one async all-reduce over elements [5, 23)
    red._start_early_bucket((red._views[2], red._views[1]))
    assert log == [("all_reduce", base + 5 * esz, 18, True, dist.ReduceOp.AVG)]
    assert red._early is not None and red._early[1:] == (5, 23)
    # a second hook call in the same step does not start a second bucket
    red._start_early_bucket((red._views[1],))
    assert len(log) == 1

    log.clear()
    red.reduce()
    ends = sorted((e[1], e[2]) for e in log if e[0] == "all_reduce")
    assert ends == [(base, 5), (base + 23 * esz, 3)]          # [0, 5) and [23, 26)
    assert all(not e[3] for e in log if e[0] == "all_reduce")
    assert log[-1] == ("wait",) and red._early is None
    assert red.collectives == 2                               # the bucket + the grouped ends
    for p, v in zip(params, red._views):
        assert p.grad.data_ptr() == v.data_ptr()

    # next step without the hook firing: a single all-reduce over the whole buffer
    log.clear()
    red.reduce()
    assert log == [("all_reduce", base, 26, False, dist.ReduceOp.AVG)]
