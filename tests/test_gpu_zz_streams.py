"""Multi-stream paths of the extractor (collected last): uploads on the copy stream recycle their buffers from batch to batch, so
every result is compared with the single-stream, device-resident path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_run_table_from_pinned_host_volumes_equals_resident_volumes(cuda):
    """PointCloudExtractor.run_table over several backbone batches whose volumes arrive from pinned host memory (uploaded on the
    copy stream while the previous batch is in the backbone; the upload buffers are recycled from batch to batch and each
    upload waits for the forward that last read its buffer) writes the table the device-resident volumes give, bit for bit."""
    from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200.distributed import PointCloudTable
    n = 6
    vols = [synth.make_case("C1", seed=70 + i) for i in range(n)]
    img, mask, res, name = vols[0]
    model = tdd.load_model(name, img_hw=img.shape[:2], device=cuda, seed=3)
    ex = tdd.PointCloudExtractor(model)
    masks_u8 = [np.ascontiguousarray(v[1]).view(np.uint8) for v in vols]
    resident = [(i, torch.as_tensor(v[0]).to(cuda), torch.as_tensor(m).to(cuda), v[2]) for i, (v, m) in enumerate(zip(vols, masks_u8))]
    pinned = [(i, torch.as_tensor(v[0]).pin_memory(), torch.as_tensor(m).pin_memory(), v[2]) for i, (v, m) in enumerate(zip(vols, masks_u8))]
    cap = n * img.shape[2] * model.grid[0] * model.grid[1]
    tables = []
    for items in (resident, pinned, pinned):
        tb = PointCloudTable(n, model.cfg["dim"], cap_rows=cap, device=cuda, rank=0, world=1)
        total = ex.run_table(items, tb, max_rows_per_batch=0)          # one patient per backbone batch: n batches
        torch.cuda.synchronize()
        tables.append((total, tb.tokens[:total].clone(), tb.src[:total].clone()))
    assert tables[0][0] > 60
    for t in tables[1:]:
        assert t[0] == tables[0][0]
        assert torch.equal(t[2], tables[0][2])
        assert torch.equal(t[1], tables[0][1])
    per_patient = torch.bincount(tables[0][2][:, 0].long(), minlength=n)
    assert int(per_patient.min()) > 0                                  # every patient contributed rows, in patient order
    assert bool((tables[0][2][1:, 0] >= tables[0][2][:-1, 0]).all())


def test_streaming_extractor_with_changing_volume_shapes_equals_one_patient_at_a_time(cuda):
    """PointCloudExtractor.run(to_host=False) over patients whose slice count changes from one to the next (as on real data): the
    upload slots are re-allocated on the copy stream while the previous patient's kernels are still queued; every point cloud
    must equal the one the same patient gives alone."""
    from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd
    cases = []
    for i in range(5):
        img, mask, res, name = synth.make_case("C1", seed=80 + i)
        s0, s1 = ((0, 8), (1, 7), (0, 7), (1, 8), (2, 7))[i]
        cases.append((np.ascontiguousarray(img[:, :, s0:s1]), np.ascontiguousarray(mask[:, :, s0:s1]), res))
    model = tdd.load_model(name, img_hw=cases[0][0].shape[:2], device=cuda, seed=3)
    ex = tdd.PointCloudExtractor(model)
    streamed = []
    for out in ex.run(cases, to_host=False):
        streamed.append((out["count"], out["tokens"], out["src"]))         # consumed later: the kernels stay queued
    torch.cuda.synchronize()
    assert len(streamed) == len(cases)
    for (count, tokens, src), (img, mask, res) in zip(streamed, cases):
        n = int(count.item())
        alone = tdd.extract_point_cloud(model, img, mask, res, to_host=False)
        m = int(alone["count"].item())
        assert n == m and n > 0
        assert torch.equal(src[:n], alone["src"][:m])
        assert torch.equal(tokens[:n], alone["tokens"][:m])
