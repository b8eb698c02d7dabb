"""GPU, >= 2 devices, NCCL: rank-averaged gradients of the sharded batch == single-process gradients of
the whole batch (OPNet), one collective per step, identical flat buffers on every rank."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

OPNET_CFG = {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 256, "videos_hidden_dim": 512}
B_GLOBAL, T = 16, 40


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grads(model, boxes, labels, dev):
    from objectpermanence_b200 import ops
    for p in model.parameters():
        p.grad = None
    y, _ = model(boxes.to(dev))
    ops.training_loss(y, labels.to(dev), None, False)[0].backward()


def _make(seed=0):
    from objectpermanence_b200.models_factory import ModelsFactory
    from objectpermanence_b200.synthetic import make_batch
    from oracle import opnet_oracle as oracle
    params = oracle.init_params("opnet", OPNET_CFG, seed=seed, scale=2.0)
    model = ModelsFactory.get_model("opnet", OPNET_CFG)
    model.load_state_dict(params)
    b, l, _ = make_batch(B_GLOBAL, T, 6, seed=77)
    return model, torch.from_numpy(b), torch.from_numpy(l)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from objectpermanence_b200.data_parallel import FlatGradAllReducer, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    dev = torch.device("cuda", rank)
    model, boxes, labels = _make()
    model = model.to(dev)
    lo, hi = shard_bounds(B_GLOBAL, rank, world)
    _grads(model, boxes[lo:hi], labels[lo:hi], dev)
    reducer = FlatGradAllReducer(model.parameters())
    flat = reducer.reduce()
    torch.cuda.synchronize()
    assert reducer.collectives == 1
    torch.save(flat.cpu(), os.path.join(out_dir, f"flat{rank}.pt"))
    # second step with the reducer in place: the weight gradients land in the flat buffer directly (no concatenation);
    # only the transposed dW_pred product is copied into its slot
    copies = reducer.copies
    _grads(model, boxes[lo:hi], labels[lo:hi], dev)
    landed = sum(p.grad.data_ptr() == v.data_ptr() for p, v in zip(reducer.params, reducer._views))
    flat2 = reducer.reduce().clone()
    torch.cuda.synchronize()
    assert landed >= len(reducer.params) - 1 and reducer.copies - copies <= 1, (landed, reducer.copies - copies)
    assert torch.equal(flat2.cpu(), torch.load(os.path.join(out_dir, f"flat{rank}.pt")))
    dist.destroy_process_group()


def test_two_rank_gradients_match_single_process(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    f0, f1 = torch.load(tmp_path / "flat0.pt"), torch.load(tmp_path / "flat1.pt")
    assert torch.equal(f0, f1)
    dev = torch.device("cuda:0")
    model, boxes, labels = _make()
    model = model.to(dev)
    _grads(model, boxes, labels, dev)
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).cpu()
    assert (f0 - want).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item())
