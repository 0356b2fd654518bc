"""Input pipeline: the on-device augmentation against torchvision's own functional ops (what the reference's
transforms compose, src/datasets.py:107-118,281-286), and the double-buffered device loader."""
import numpy as np
import pytest
import torch


def _tv_reference(x8, params, res, pad_top, pad_left):
    import torchvision.transforms.functional as TF
    out = []
    for i in range(x8.shape[0]):
        t, l, f = (int(v) for v in params[i])
        img = TF.pad(x8[i], [pad_left, pad_top], fill=0)          # RandomCrop(padding=[left/right, top/bottom])
        img = TF.crop(img, t, l, res, res)
        out.append(TF.hflip(img) if f else img)
    return torch.stack(out)


def test_augment_checker_equals_torchvision_ops():
    from causalgen_b200.pipeline import Augment
    g = torch.Generator().manual_seed(1)
    for (n, c, hi, wi, res, pt, pl, hf) in [(5, 1, 28, 28, 32, 4, 4, 0.0), (4, 1, 40, 40, 40, 3, 6, 0.5), (3, 3, 32, 32, 32, 4, 4, 0.5)]:
        x8 = torch.randint(0, 256, (n, c, hi, wi), generator=g, dtype=torch.uint8)
        aug = Augment(res, pt, pl, hf, seed=3)
        prm = aug.draw(n, hi, wi)
        assert prm.dtype == torch.int32 and prm.shape == (n, 3)
        assert int(prm[:, 0].max()) <= hi + 2 * pt - res and int(prm[:, 1].max()) <= wi + 2 * pl - res and int(prm.min()) >= 0
        assert torch.equal(Augment.reference(x8, prm, res, pt, pl), _tv_reference(x8, prm, res, pt, pl))


@pytest.mark.gpu
def test_augment_kernel_bit_exact_and_device_loader_order():
    from causalgen_b200.pipeline import Augment, DeviceLoader
    g = torch.Generator().manual_seed(2)
    for (n, c, hi, wi, res, pt, pl, hf) in [(7, 1, 28, 28, 32, 4, 4, 0.0), (5, 1, 192, 192, 192, 9, 18, 0.5), (6, 3, 32, 32, 32, 4, 4, 0.5),
                                             (2, 1, 1, 1, 1, 0, 0, 1.0)]:
        x8 = torch.randint(0, 256, (n, c, hi, wi), generator=g, dtype=torch.uint8)
        aug = Augment(res, pt, pl, hf, seed=5)
        prm = aug.draw(n, hi, wi)
        prm[0] = torch.tensor([0, 0, int(hf > 0)])                                     # corner cases: extreme offsets
        prm[-1] = torch.tensor([hi + 2 * pt - res, wi + 2 * pl - res, 0])
        got = aug.apply(x8.cuda(), prm.cuda()).cpu()
        assert torch.equal(got, Augment.reference(x8, prm, res, pt, pl)), (n, c, hi, wi)
    # loader: every batch arrives once, in order, bit-identical (no augmentation), while copies overlap the consumer
    batches = [(torch.randint(0, 256, (4, 1, 16, 16), generator=g, dtype=torch.uint8), torch.randn(4, 3, generator=g)) for _ in range(7)]
    seen = []
    for xd, pd in DeviceLoader(batches):
        seen.append((xd.clone(), pd.clone()))
        torch.cuda._sleep(200000)  # consumer busy: the next copy must not overwrite what was handed out
    torch.cuda.synchronize()
    assert len(seen) == len(batches)
    for (xd, pd), (x, p) in zip(seen, batches):
        assert torch.equal(xd.cpu(), x) and torch.equal(pd.cpu(), p)
    # with augmentation the loader output equals the checker on the params it drew (same generator seed)
    aug = Augment(16, 2, 2, 0.5, seed=9)
    ref_aug = Augment(16, 2, 2, 0.5, seed=9)
    for (xd, pd), (x, p) in zip(DeviceLoader(batches, augment=aug), batches):
        prm = ref_aug.draw(x.shape[0], 16, 16)
        assert torch.equal(xd.cpu(), Augment.reference(x, prm, 16, 2, 2))
