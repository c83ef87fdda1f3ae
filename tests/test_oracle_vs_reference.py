"""Pins the oracle restatement against the unmodified reference modules imported from /root/reference
(build container only; skipped on the GPU box where the tree is absent)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import cases, nets, refload, weights

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not refload.available(), reason="/root/reference not present")]
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def ref():
    warnings.filterwarnings("ignore")
    return refload.load()


def _check_keys(module, sd):
    want = {k: tuple(v.shape) for k, v in module.state_dict().items()
            if "num_batches_tracked" not in k and not k.endswith("pe.pe")}
    have = {k: tuple(v.shape) for k, v in sd.items()}
    assert want == have


def test_state_dict_specs_match_reference(ref):
    _check_keys(ref["det"].DBNetConvNext(), weights.dbnet_weights())
    _check_keys(ref["ocr"].OCR(["x"] * 300, 768), weights.ocr_weights(300))
    for nb in (9, 18):
        lf = ref["lama"].LamaFourier(build_discriminator=False, use_mpe=nb == 9, large_arch=nb == 18)
        _check_keys(lf.generator, weights.lama_weights(nb))
        if nb == 9:
            _check_keys(lf.mpe, weights.mpe_weights())
            assert torch.equal(lf.mpe.rel_pos_emb.weight, weights.mpe_weights()["rel_pos_emb.weight"])


def test_dbnet_rectangular(ref):
    sd = weights.dbnet_weights(seed=2)
    net = ref["det"].DBNetConvNext().eval()
    net.load_state_dict(sd)
    _, x = cases.dbnet_case(256, 512, seed=21)
    r_db, r_mask = net(x)
    o_db, o_mask = nets.dbnet_forward(sd, x)
    assert (r_db - o_db).abs().max() < 1e-4 and (r_mask - o_mask).abs().max() < 1e-5


def test_ocr_widths_and_decode(ref):
    V = 300
    sd = weights.ocr_weights(V, seed=3)
    ocr = ref["ocr"].OCR(weights.synthetic_dictionary(V), 768).eval()
    ocr.load_state_dict(sd, strict=False)
    for wp in (143, 200, 331):
        _, x = cases.ocr_case(3, wp, seed=wp)
        rl, rc = ocr(x)
        ol, oc = nets.ocr_forward(sd, x)
        assert (rl - ol).abs().max() < 1e-4 and (rc - oc).abs().max() < 1e-5
        idx, lp, col = nets.ocr_top1(sd, x)
        ref_dec = ocr.decode(x, [0] * 3, 0)
        mine = nets.ctc_greedy(idx.numpy(), lp.numpy(), col.numpy())
        top2 = rl.topk(2, dim=-1).values
        if (top2[..., 0] - top2[..., 1]).min() > 1e-3:
            assert [[int(c[0]) for c in l] for l in ref_dec] == [[c[0] for c in l] for l in mine]


def test_lama_mpe_tables_random_masks(ref):
    lf = ref["lama"].LamaFourier(build_discriminator=False, use_mpe=True)
    rng = np.random.default_rng(5)
    for (h, w) in ((256, 256), (200, 312), (64, 48)):
        m = np.zeros((h, w), np.float32)
        for _ in range(4):
            y, x = rng.integers(0, h - 8), rng.integers(0, w - 8)
            m[y:y + rng.integers(4, h // 2), x:x + rng.integers(4, w // 2)] = 1
        rel, _, direct = lf.load_masked_position_encoding(m)
        orel, odirect = nets.mpe_tables(m)
        assert np.array_equal(rel, orel) and np.array_equal(direct, odirect)
    # all-hole and no-hole masks terminate (the reference guards the infinite loop, :778)
    for m in (np.zeros((64, 64), np.float32), np.ones((64, 64), np.float32)):
        rel, _, direct = lf.load_masked_position_encoding(m)
        orel, odirect = nets.mpe_tables(m)
        assert np.array_equal(rel, orel) and np.array_equal(direct, odirect)


def test_lama_odd_spectrum_sizes(ref):
    sd, msd = weights.lama_weights(9, seed=4), weights.mpe_weights(seed=4)
    lf = ref["lama"].LamaFourier(build_discriminator=False, use_mpe=True)
    lf.generator.load_state_dict(sd)
    lf.mpe.load_state_dict(msd)
    lf.eval()
    img, mask = cases.lama_case(88, 120, seed=41)   # bottleneck 11x15: odd FFT lengths
    r = lf(img.clone(), mask)
    rel, direct = nets.mpe_tables(mask[0, 0].numpy())
    o = nets.lama_forward(sd, msd, img, mask, torch.from_numpy(rel)[None], torch.from_numpy(direct)[None])
    assert (r - o).abs().max() < 2e-5
