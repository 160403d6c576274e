"""The N>1 host logic on CPU: world_size-2 gloo processes (all_gather_table, allreduce_grads, sharding)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vit_deep_radiomics_b200.distributed import all_gather_table, allreduce_grads, shard_modulo, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    keys = torch.stack([torch.arange(40) // 10, torch.randperm(40, generator=g) % 7, torch.arange(40) % 5], 1).int()
    rows = torch.randn(40, 6, generator=g)
    mine = shard_modulo(40, rank, world)
    k, r = all_gather_table(keys[mine], rows[mine])
    # canonical order == single-process sort of the full table
    order = torch.arange(40)
    for col in (2, 1, 0):
        order = order[torch.sort(keys[order, col], stable=True).indices]
    ok_table = torch.equal(k, keys[order]) and torch.equal(r, rows[order])
    # contiguous slice sharding: rank-order concatenation already sorted
    lo, hi = shard_range(40, rank, world)
    k2, r2 = all_gather_table(keys[lo:hi], rows[lo:hi], sort=False)
    ok_contig = torch.equal(k2, keys) and torch.equal(r2, rows)
    # gradient all-reduce == accumulation over all samples on one process
    torch.manual_seed(1)
    model = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Linear(4, 2))
    data = torch.randn(8, 6, generator=g)
    for i in range(rank, 8, world):
        (model(data[i]).sum() / 8).backward()
    allreduce_grads(model)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Linear(4, 2))
    ref.load_state_dict(model.state_dict())
    for i in range(8):
        (ref(data[i]).sum() / 8).backward()
    ok_grad = all(torch.allclose(a.grad, b.grad, atol=1e-6) for a, b in zip(model.parameters(), ref.parameters()))
    out[rank] = (ok_table, ok_contig, ok_grad)
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        assert out[rank] == (True, True, True), (rank, out[rank])


def test_shards_partition():
    for n in (1, 7, 120, 1000):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert sorted(sum((shard_modulo(n, r, w) for r in range(w)), [])) == list(range(n))


def _dp_worker(rank, world, port, out):
    """Data-parallel run_epoch (world size 2, gloo, CPU stub classifier + oracle gather) == the single-process epoch."""
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gather_np as G
    from oracle import ref_shim
    from vit_deep_radiomics_b200 import train_models as tm
    D = 12
    df = tm.prepare_df(ref_shim.make_dataset_table(seed=4, D=D))
    enc = tm.get_label_encoder(df)
    ds = tm.PETCTDataset3D(df, enc, "ct.h5", "pet.h5", use_augmentation=False, feature_dim=D, arch="transformer", store=ref_shim.H5_FILES,
                           gather=lambda f, m, r, n, d: G.token_gather(f, m, r, n, d)["tokens"])

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(D, 2)

        def forward(self, x):
            cls = x.mean(1)
            return self.lin(cls), cls

    def epoch(r, w, sync):
        torch.manual_seed(5)
        model = Stub()
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        crit = tm.FocalLoss(alpha=torch.tensor([0.25, 0.75]), gamma=2.0)
        order = torch.randperm(len(ds), generator=torch.Generator().manual_seed(1)).tolist()
        res = tm.run_epoch(model, ds, order, crit, "ct", "cpu", opt, virtual_batch_size=4, grad_sync=sync, rank=r, world=w)
        ev = tm.run_epoch(model, ds, range(len(ds)), crit, "ct", "cpu", rank=r, world=w)
        return model, res, ev

    m_dp, res_dp, ev_dp = epoch(rank, world, allreduce_grads)
    m_1, res_1, ev_1 = epoch(0, 1, None)
    ok_w = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(m_dp.state_dict().values(), m_1.state_dict().values()))
    ok_loss = abs(res_dp[0] - res_1[0]) < 1e-6 and abs(ev_dp[0] - ev_1[0]) < 1e-6
    ok_n = len(res_dp[1]) == len(res_1[1]) == len(ds) and sorted(np.concatenate(ev_dp[3]).tolist()) == sorted(np.concatenate(ev_1[3]).tolist())
    out[rank] = (ok_w, ok_loss, ok_n, len(ds))
    dist.destroy_process_group()


def test_data_parallel_epoch_equals_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        assert out[rank][:3] == (True, True, True) and out[rank][3] > 4, (rank, out[rank])


def _fold_worker(rank, world, port, out, save_dir):
    """run_fold at world size 2 (gloo, CPU stub classifier): rank 0's weights / epoch orders are broadcast, every rank sees the
    gathered records (identical histories), rank 0 alone writes the metric files and checkpoints."""
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gather_np as G
    from oracle import ref_shim
    from vit_deep_radiomics_b200 import config_manager, train_models as tm
    D = 12
    df = tm.prepare_df(ref_shim.make_dataset_table(seed=6, D=D))
    enc = tm.get_label_encoder(df)
    cfg = config_manager.load_conf(project_dir=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cfg["models"]["transformer"].update(feature_dim=D, patience=5, virtual_batch_size=4)

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(D, 2)

        def forward(self, x):
            cls = x.mean(1)
            return self.lin(cls), cls

    tm.build_model = lambda *a, **k: Stub()           # each rank initialises differently (seed below): the broadcast must align them
    torch.manual_seed(100 + rank)
    np.random.seed(200 + rank)
    my_dir = os.path.join(save_dir, f"rank{rank}")
    hist = tm.run_fold(cfg, "transformer", "ct", df[df.patient_id.isin(["P1", "P2"])].reset_index(drop=True),
                       df[df.patient_id.isin(["P3", "P4"])].reset_index(drop=True), enc, "ct.h5", "pet.h5", my_dir, kfold=1,
                       device="cpu", store=ref_shim.H5_FILES, num_epochs=2, rank=rank, world=world,
                       gather=lambda f, m, r, n, d: G.token_gather(f, m, r, n, d)["tokens"])
    out[rank] = ([(h["epoch"], round(h["train_loss"], 9), round(h["test_loss"], 9), round(h["test_auc"], 9)) for h in hist],
                 sorted(os.listdir(my_dir)) if os.path.isdir(my_dir) else None)
    dist.destroy_process_group()


def test_run_fold_world_size_2(tmp_path):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fold_worker, args=(world, _free_port(), out, str(tmp_path)), nprocs=world, join=True)
    assert out[0][0] == out[1][0] and len(out[0][0]) == 2              # same gathered records -> same history on both ranks
    assert out[1][1] is None                                           # rank 1 writes nothing
    assert {"train_metrics_0.json", "test_metrics_0.json", "train_metrics_1.json", "test_metrics_1.json", "model_epoch_0000.pth"} <= set(out[0][1])


def _table_worker(rank, world, port, out, n_rows=(4, 0, 7, 3, 5)):
    """PointCloudTable host logic at world size 2 (gloo, CPU tensors): contiguous patient shards, padded count exchange,
    offsets, in-place variable-length all-gather == the single-process table, on every rank.  (The device side -- the gather
    kernel writing at a device-resident offset -- is covered by tests/test_gpu_gather.py::test_g1_count_scan_and_table_slots.)"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vit_deep_radiomics_b200.distributed import PointCloudTable
    g = torch.Generator().manual_seed(3)
    P, D = 5, 6                                                   # 5 patients over 2 ranks: shards of 3 and 2 (padded to 3)
    n_rows = list(n_rows)
    clouds = [torch.randn(n, D, generator=g) for n in n_rows]
    keys = [torch.stack([torch.full((n,), p), torch.arange(n) % 3, torch.arange(n) // 3, torch.arange(n)], 1).int() for p, n in enumerate(n_rows)]
    table = PointCloudTable(P, D, cap_rows=sum(n_rows) + 2, device="cpu")
    for p in table.local_patients():
        table.count_out(p)[0] = n_rows[p]
    table.exchange_counts()
    for p in table.local_patients():                              # what vdr_mask_gather_table does on the device
        slot = table.slot(p)
        off = int(slot["row_offset"][0])
        slot["tokens"][off:off + n_rows[p]] = clouds[p]
        slot["src"][off:off + n_rows[p]] = keys[p]
    total = table.all_gather()
    ok = total == sum(n_rows) and torch.equal(table.tokens[:total], torch.cat(clouds)) and torch.equal(table.src[:total], torch.cat(keys))
    # padded, rank-major count vector: rank 0 = patients 0..2, rank 1 = patients 3, 4 + one pad slot
    padded = n_rows[:3] + n_rows[3:] + [0]
    ok_off = table.offsets.tolist() == [0] + [sum(padded[:i + 1]) for i in range(6)]
    out[rank] = (ok, ok_off, list(table.local_patients()))
    dist.destroy_process_group()


def test_point_cloud_table_world_size_2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_table_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == (True, True, [0, 1, 2]) and out[1] == (True, True, [3, 4]), dict(out)


def test_point_cloud_table_rank_without_rows():
    """A rank whose patients select nothing (every mask empty inside its ROI) contributes an empty range: the exchange skips it
    and both ranks still end with the single-process table."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_table_worker, args=(world, _free_port(), out, (4, 2, 7, 0, 0)), nprocs=world, join=True)
    assert out[0] == (True, True, [0, 1, 2]) and out[1] == (True, True, [3, 4]), dict(out)


def _bucket_worker(rank, world, port, out):
    """GradBucket: .grad tensors are views of one flat buffer; all-reduce through it == per-tensor all-reduce, also after
    ``zero_grad(set_to_none=True)`` detached the views."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vit_deep_radiomics_b200.distributed import grad_bucket, zero_grads
    torch.manual_seed(1)
    model = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Linear(4, 2))
    data = torch.randn(8, 6, generator=torch.Generator().manual_seed(0))
    ref = [torch.zeros_like(p) for p in model.parameters()]
    for i in range(8):
        model.zero_grad()
        model(data[i]).sum().backward()
        for r, p in zip(ref, model.parameters()):
            r += p.grad
    oks = []
    for mode in ("views", "detached"):
        if mode == "views":
            zero_grads(model)
            grad_bucket(model).zero()
        else:
            model.zero_grad(set_to_none=True)
        for i in range(rank, 8, world):
            model(data[i]).sum().backward()
        allreduce_grads(model)
        b = grad_bucket(model)
        oks.append(all(torch.allclose(p.grad, r, atol=1e-6) for p, r in zip(model.parameters(), ref)))
        oks.append(all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(b.params, b.views)))
    out[rank] = tuple(oks)
    dist.destroy_process_group()


def test_grad_bucket_world_size_2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == (True,) * 4 and out[1] == (True,) * 4, dict(out)
