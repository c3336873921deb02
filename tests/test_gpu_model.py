"""The WholeGraph-native training path (reference: python/pylibwholegraph/examples/node_classfication.py:124-200):
GraphStructure layered sampling -> WholeMemoryEmbeddingModule gather -> HomoGNNModel (aggregation on the sampler's CSR
blocks) -> loss, with optional trainable node embeddings."""
import argparse

import numpy as np
import pytest

from graphs import random_csr

pytestmark = pytest.mark.gpu


def _args(model, train_embedding=False):
    import pylibwholegraph.torch as wgth

    p = argparse.ArgumentParser()
    for f in (wgth.add_training_options, wgth.add_common_graph_options, wgth.add_common_model_options,
              wgth.add_common_sampler_options, wgth.add_node_classfication_options, wgth.add_dataloader_options):
        f(p)
    argv = ["--model", model, "--layernum", "2", "--hiddensize", "32", "--neighbors", "10,5", "--inferencesample", "10",
            "--classnum", "4", "--dropout", "0.0", "--framework", "wg"] + (["--train-embedding"] if train_embedding else [])
    return p.parse_args(argv)


@pytest.mark.parametrize("model,train_embedding", [("sage", False), ("gcn", False), ("sage", True)])
def test_homo_gnn_model_learns_planted_labels(model, train_embedding):
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    comm = wgth.get_global_communicator()
    torch.manual_seed(0)
    nodes, edges, dim = 4000, 60000, 16
    row_ptr, col = random_csr(nodes, edges, seed=5)
    rng = np.random.default_rng(0)
    labels = rng.integers(0, 4, nodes)
    # neighbours mostly share the label of their centre: rewire columns to same-label vertices
    by_label = [np.nonzero(labels == c)[0] for c in range(4)]
    centre = np.repeat(np.arange(nodes), np.diff(row_ptr))
    same = rng.random(edges) < 0.8
    col = np.where(same, np.array([by_label[labels[c]][rng.integers(len(by_label[labels[c]]))] for c in centre]), col).astype(np.int32)
    feat = (np.eye(4, dim)[labels] + 0.5 * rng.standard_normal((nodes, dim))).astype(np.float32)

    def wm(arr):
        t = torch.from_numpy(arr)
        w = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [arr.shape[0]], t.dtype, [1])
        w.get_local_tensor()[0].copy_(t.cuda())
        return w

    gs = wgth.GraphStructure()
    gs.set_csr_graph(wm(row_ptr), wm(col))
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [nodes, dim])
    emb.get_embedding_tensor().get_local_tensor()[0].copy_(torch.from_numpy(feat if not train_embedding else 0.01 * feat).cuda())
    args = _args(model, train_embedding)
    wm_opt = wgth.create_wholememory_optimizer(emb, "adam", {}) if train_embedding else None
    net = wgth.HomoGNNModel(gs, emb, args).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=0.02)
    y = torch.from_numpy(labels).cuda()
    train_sets, _, _ = wgth.create_node_classification_datasets({
        "train_idx": np.arange(3000), "train_label": labels[:3000], "valid_idx": np.arange(3000, 3500),
        "valid_label": labels[3000:3500], "test_idx": np.arange(3500, 4000), "test_label": labels[3500:]})
    loader = wgth.get_train_dataloader(train_sets, 256)
    losses = []
    net.train()
    for epoch in range(4):
        for ids, lab in loader:
            out = net(ids)
            loss = torch.nn.functional.cross_entropy(out, lab.cuda())
            opt.zero_grad()
            loss.backward()
            opt.step()
            if wm_opt is not None:
                wm_opt.step(0.05)
            losses.append(float(loss.detach()))
    net.eval()
    with torch.no_grad():
        test_ids = torch.arange(3500, 4000)
        acc = float((net(test_ids).argmax(1) == y[3500:]).float().mean())
    assert losses[-1] < 0.7 * losses[0], (losses[0], losses[-1])
    assert acc > 0.6, acc  # 4 classes: chance is 0.25
