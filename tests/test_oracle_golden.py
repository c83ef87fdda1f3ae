"""The oracle restatement against the committed outputs of the reference modules (tests/golden, made by
oracle/make_golden.py).  CPU only.  Tolerances: the restatement uses the same torch CPU kernels as the
reference, so agreement is at fp32 round-off (<=2e-5 observed); strings / indices must be identical."""
import os

import numpy as np
import torch

from oracle import cases, nets, weights

torch.set_grad_enabled(False)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_dbnet_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "dbnet_256.npz")
    _, x = cases.dbnet_case()
    db, mask = nets.dbnet_forward(weights.dbnet_weights(), x)
    assert np.abs(db.sigmoid().numpy() - g["db_sigmoid"]).max() < 2e-5
    assert np.abs(mask.numpy() - g["mask"]).max() < 2e-5
    assert np.abs(db[:, 0].numpy() - g["db_logit0"]).max() < 2e-4


def test_dbnet_batch_forward_normalisation():
    # divide-then-subtract (dbnet_convnext.py:503) is NOT bitwise (x-127.5)/127.5; keep the op order
    img, x = cases.dbnet_case(256, 256)
    a = img.astype(np.float32) / 127.5 - 1.0
    assert np.array_equal(a.transpose(0, 3, 1, 2), x.numpy())


def test_ocr_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "ocr_200.npz")
    _, x = cases.ocr_case()
    idx, lp, col = nets.ocr_top1(weights.ocr_weights(cases.OCR_VOCAB_SMALL), x)
    safe = g["margin"] > 1e-3
    assert np.array_equal(idx.numpy()[safe], g["idx"][safe])
    assert np.abs(lp.numpy() - g["logprob"])[safe].max() < 1e-4
    assert np.abs(col.numpy() - g["colors"]).max() < 1e-4
    dec = nets.ctc_greedy(idx.numpy(), lp.numpy(), col.numpy())
    flat = [(b, c[0]) for b, line in enumerate(dec) for c in line]
    assert flat == [(int(r[0]), int(r[1])) for r in g["decoded"]]
    assert idx.shape[1] == (x.shape[-1] // 2) // 2 - 1  # T = floor(floor(Wp/2)/2) - 1


def test_ctc_greedy_handcrafted():
    idx = np.array([[0, 5, 5, 0, 5, 7, 7, 7, 0, 0, 3]])
    lp = -np.arange(11, dtype=np.float32)[None] / 10
    col = np.zeros((1, 11, 6), np.float32)
    out = nets.ctc_greedy(idx, lp, col)
    assert [c[0] for c in out[0]] == [5, 5, 7, 3]
    assert [round(c[1], 3) for c in out[0]] == [-0.1, -0.4, -0.5, -1.0]


def test_lama_matches_reference_fixture(golden_dir):
    img, mask = cases.lama_case()
    g = _load(golden_dir, "lama_mpe_128x96.npz")
    rel, direct = nets.mpe_tables(mask[0, 0].numpy())
    assert np.array_equal(rel, g["rel_pos"].astype(np.int32))
    assert np.array_equal(direct, g["direct"].astype(np.int32))
    assert rel.min() >= 0 and rel.max() <= 127
    out = nets.lama_forward(weights.lama_weights(9), weights.mpe_weights(), img, mask,
                            torch.from_numpy(rel)[None], torch.from_numpy(direct)[None])
    assert np.abs(out.numpy() - g["out"]).max() < 2e-5
    g = _load(golden_dir, "lama_large_128x96.npz")
    out = nets.lama_forward(weights.lama_weights(18), None, img, mask)
    assert np.abs(out.numpy() - g["out"]).max() < 2e-5


def test_ffc_block_and_fourier_unit_fixture(golden_dir):
    g = _load(golden_dir, "ffc_block_20x14.npz")
    sd = weights.lama_weights(1)
    rng = np.random.default_rng(14)
    xl = torch.from_numpy(rng.standard_normal((1, 128, 20, 14)).astype(np.float32))
    xg = torch.from_numpy(rng.standard_normal((1, 384, 20, 14)).astype(np.float32))
    yl, yg = nets.ffc_bn_act(sd, "model.5.conv1.", xl, xg)
    yl, yg = nets.ffc_bn_act(sd, "model.5.conv2.", yl, yg)
    assert np.abs((xl + yl).numpy() - g["yl"]).max() < 2e-5
    assert np.abs((xg + yg).numpy() - g["yg"]).max() < 2e-5
    s = torch.from_numpy(rng.standard_normal((1, 192, 20, 14)).astype(np.float32))
    fu = nets.fourier_unit(sd, "model.5.conv1.ffc.convg2g.fu.", s)
    assert np.abs(fu.numpy() - g["fu"]).max() < 2e-5


def test_fourier_unit_closed_form():
    # identity spectral conv, BN = identity, positive spectrum => irfft(rfft(x)) == x
    c = 4
    sd = {"p.conv_layer.weight": torch.eye(2 * c).reshape(2 * c, 2 * c, 1, 1),
          "p.bn.weight": torch.ones(2 * c) * (1 + 1e-5) ** 0.5, "p.bn.bias": torch.zeros(2 * c),
          "p.bn.running_mean": torch.zeros(2 * c), "p.bn.running_var": torch.ones(2 * c)}
    x = torch.zeros(1, c, 6, 10)
    x[:, :, 0, 0] = 3.0  # impulse: spectrum is constant, real and positive -> ReLU is a no-op
    assert (nets.fourier_unit(sd, "p.", x) - x).abs().max() < 1e-5


def test_mpe_single_rectangle():
    m = np.zeros((256, 256), np.float32)
    m[100:140, 60:200] = 1
    rel, direct = nets.mpe_tables(m)
    assert rel[0, 0] == 0 and rel[120, 130] == 20  # centre row is 20 dilations from the top edge (rows 100..139)
    assert rel[100, 130] == 1 and direct[~(m > 0)].sum() == 0
