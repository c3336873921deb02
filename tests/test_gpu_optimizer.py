"""(f.4) Trainable WholeMemory embeddings: wholememory_embedding_gather_gradient_apply with the four sparse optimizers
against the oracle's restatement of the reference's update rules (embedding_optimizer_func.cu), including repeated
indices, several steps (per-row Adam bias correction), fp16/bf16 tables, and the torch autograd path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    ("sgd", {"weight_decay": 0.01}),
    ("adam", {"weight_decay": 0.02, "epsilon": 1e-6, "beta1": 0.8, "beta2": 0.95, "adam_w": 0.0}),
    ("adam", {"weight_decay": 0.02, "adam_w": 1.0}),
    ("adagrad", {"weight_decay": 0.0, "epsilon": 1e-5}),
    ("rmsprop", {"weight_decay": 0.01, "epsilon": 1e-6, "alpha": 0.9}),
]


@pytest.fixture(scope="module")
def env():
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    return torch, wgth, wgth.get_global_communicator()


def _states(name, rows, dim):
    if name == "adam":
        return {"m": np.zeros((rows, dim), np.float32), "v": np.zeros((rows, dim), np.float32), "beta12t": np.ones((rows, 2), np.float32)}
    if name == "adagrad":
        return {"state_sum": np.zeros((rows, dim), np.float32)}
    if name == "rmsprop":
        return {"v": np.zeros((rows, dim), np.float32)}
    return {}


@pytest.mark.parametrize("opt,params", CASES)
@pytest.mark.parametrize("dim,idx_dtype", [(128, np.int64), (37, np.int32)])
def test_gradient_apply_matches_oracle(env, oracle, opt, params, dim, idx_dtype):
    torch, wgth, comm = env
    rows = 5003
    rng = np.random.default_rng(dim)
    table = rng.standard_normal((rows, dim)).astype(np.float32)
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [rows, dim])
    emb.get_embedding_tensor().get_local_tensor()[0].copy_(torch.from_numpy(table).cuda())
    optimizer = wgth.create_wholememory_optimizer(emb, opt, params)
    assert emb.get_optimizer_state_names() == list(_states(opt, 1, 1).keys())
    ref, states = table.copy(), _states(opt, rows, dim)
    for step in range(3):
        before = emb.get_embedding_tensor().get_local_tensor()[0].cpu().numpy()
        n = 4000
        idx = rng.integers(0, rows, n)
        idx[: n // 4] = rng.integers(0, 40, n // 4)  # heavy repetition: hubs
        grads = rng.standard_normal((n, dim)).astype(np.float32)
        lr = 0.05 / (step + 1)
        emb.add_gradients(torch.from_numpy(idx.astype(idx_dtype)).cuda(), torch.from_numpy(grads).cuda())
        emb.need_apply = True
        optimizer.step(lr)
        oracle.embedding_gradient_apply(opt, params, ref, idx, grads, lr, states)
        got = emb.get_embedding_tensor().get_local_tensor()[0].cpu().numpy()
        # duplicates are summed in the same order on both sides; fused multiply-adds on the device leave a few ulps
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6)
        untouched = np.setdiff1d(np.arange(rows), idx)
        assert np.array_equal(got[untouched], before[untouched])  # rows without gradients are bit-identical
    for name, exp in states.items():
        got = emb.get_optimizer_state(name).get_local_tensor()[0].cpu().numpy()
        np.testing.assert_allclose(got, exp, rtol=1e-4, atol=1e-6)  # fma contraction differs between nvcc and gcc
    wgth.destroy_wholememory_optimizer(optimizer)
    wgth.destroy_embedding(emb)


@pytest.mark.parametrize("dtype_name", ["float16", "bfloat16"])
def test_gradient_apply_half_tables(env, oracle, dtype_name):
    torch, wgth, comm = env
    dtype = getattr(torch, dtype_name)
    rows, dim = 1000, 64
    rng = np.random.default_rng(1)
    table = torch.from_numpy(rng.standard_normal((rows, dim)).astype(np.float32)).to(dtype)
    emb = wgth.create_embedding(comm, "chunked", "cuda", dtype, [rows, dim])
    emb.get_embedding_tensor().get_local_tensor()[0].copy_(table.cuda())
    optimizer = wgth.create_wholememory_optimizer(emb, "sgd", {"weight_decay": 0.0})
    idx = rng.permutation(rows)[:300]
    grads = rng.standard_normal((300, dim)).astype(np.float32)
    emb.add_gradients(torch.from_numpy(idx).cuda(), torch.from_numpy(grads).cuda())
    emb.need_apply = True
    optimizer.step(0.1)
    exp = table.float().clone()
    exp[torch.from_numpy(idx)] -= 0.1 * torch.from_numpy(grads)
    got = emb.get_embedding_tensor().get_local_tensor()[0].cpu()
    assert torch.equal(got, exp.to(dtype))  # fp32 maths, one rounding on store (as the reference's static_cast)
    wgth.destroy_embedding(emb)


def test_embedding_module_autograd_training_loop(env):
    """The reference's usage: WholeMemoryEmbeddingModule in train() mode + WholeMemoryOptimizer.step(lr)."""
    torch, wgth, comm = env
    rows, dim = 2000, 32
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [rows, dim], random_init=True)
    start = emb.get_embedding_tensor().get_local_tensor()[0].clone()
    optimizer = wgth.create_wholememory_optimizer(emb, "sgd", {})
    module = wgth.WholeMemoryEmbeddingModule(emb).train()
    idx = torch.tensor([5, 9, 5, 77, 1999], device="cuda")
    target = torch.ones((5, dim), device="cuda")
    losses = []
    for _ in range(20):
        out = module(idx)
        loss = ((out - target) ** 2).sum()
        loss.backward()
        optimizer.step(0.05)
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.05 * losses[0]
    now = emb.get_embedding_tensor().get_local_tensor()[0]
    touched = torch.zeros(rows, dtype=torch.bool, device="cuda")
    touched[idx] = True
    assert torch.equal(now[~touched], start[~touched])
    module.eval()
    before = emb.get_embedding_tensor().get_local_tensor()[0].clone()
    module(idx).sum().backward()  # eval mode: gradients are not collected
    optimizer.step(0.05)
    assert torch.equal(emb.get_embedding_tensor().get_local_tensor()[0], before)
    with pytest.raises(ValueError):
        wgth.create_wholememory_optimizer(emb, "adam", {})  # optimizer can only be set once
    with pytest.raises(ValueError):
        wgth.create_wholememory_optimizer(wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [8, 4]), "sgd", {"beta1": 0.5})
