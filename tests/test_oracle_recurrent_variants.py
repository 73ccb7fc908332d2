"""CPU tests of the oracle's restatement of the two recurrent_DCGAN variants
(models/recurrent_image/rnn_test/multi-layer_recurrent_DCGAN.py, ..._with_shared_conv_and_drop_out.py): the layer stack
against a hand-unrolled MultiRNNCell, where DropoutWrapper acts, which variables each optimiser reaches when the encoder
shares the discriminator's filters, and the committed golden fixture."""
import os
import sys

import numpy as np
import torch

from oracle import tf_ops as T
from oracle.models import RecurrentDCGAN

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
f64 = torch.float64
INP = np.random.RandomState(104).randint(0, 256, (2, 4, 64, 64, 3)).astype(np.int32)


def test_multi_layer_stack_equals_hand_unrolled_cells():
    m = RecurrentDCGAN(batch_size=2, video_length=3, dtype=f64, num_layers=3)
    m.trace = {}
    X, _ = m._split(torch.tensor(INP))
    with torch.no_grad():
        m.generator(X)
        enc = []
        for x in X:
            for i in range(4):
                x = torch.relu(T.batch_norm_plain(T.conv2d(x, m.vars[f"generator/conv_f{i+1}"])))
            enc.append(x.reshape(2, -1))
        st = [(torch.zeros(2, 100, dtype=f64), torch.zeros(2, 100, dtype=f64)) for _ in range(3)]
        for t in range(3):
            inp = enc[t]
            for k in range(3):      # tf.nn.rnn_cell.MultiRNNCell.__call__: cur_inp, new_state = cell(cur_inp, state[k])
                c, h = T.basic_lstm_cell(inp, st[k][0], st[k][1], m.vars[f"generator/lstm/Cell{k}/Matrix"], m.vars[f"generator/lstm/Cell{k}/Bias"])
                st[k] = (c, h)
                inp = h
            assert torch.allclose(m.trace[f"lstm_h{t}"], inp, rtol=0, atol=1e-14), t
    # separate parameters per layer (MultiRNNCell scopes Cell0..Cell2), layer 0 sees the 8192-wide encoding
    assert m.vars["generator/lstm/Cell0/Matrix"].shape == (8292, 400) and m.vars["generator/lstm/Cell1/Matrix"].shape == (200, 400)
    assert not torch.equal(m.vars["generator/lstm/Cell1/Matrix"], m.vars["generator/lstm/Cell2/Matrix"])


def test_dropout_scales_cell_outputs_not_the_recurrent_state():
    kw = dict(batch_size=2, video_length=3, dtype=f64, num_layers=3, shared_conv=True)
    a, b = RecurrentDCGAN(output_keep_prob=1.0, **kw), RecurrentDCGAN(output_keep_prob=0.8, **kw)
    a.trace, b.trace = {}, {}
    X, _ = a._split(torch.tensor(INP))
    with torch.no_grad():
        b.masks = torch.ones(3, 3, 2, 100, dtype=f64)                      # keep everything, scale 1: identical to no wrapper
        ya, yb = a.generator(X), b.generator(X)
        assert all(torch.equal(u, v) for u, v in zip(ya, yb))
        # drop the TOP layer's output at t = 0 only: frame 0's decoder sees zeros, later frames are unchanged because the
        # state handed to t = 1 is the undropped one
        b.masks = torch.ones(3, 3, 2, 100, dtype=f64)
        b.masks[2, 0] = 0.0
        yc = b.generator(X)
        assert float(b.trace["lstm_h0"].abs().max()) == 0.0
        assert not torch.equal(yc[0], ya[0]) and torch.equal(yc[1], ya[1]) and torch.equal(yc[2], ya[2])
        # dropping a LOWER layer's output at t = 0 changes what the upper layers store, hence every later frame
        b.masks = torch.ones(3, 3, 2, 100, dtype=f64)
        b.masks[0, 0] = 0.0
        yd = b.generator(X)
        assert not torch.equal(yd[1], ya[1])


def test_shared_encoder_gradient_paths():
    m = RecurrentDCGAN(batch_size=2, video_length=3, dtype=f64, num_layers=3, shared_conv=True, output_keep_prob=0.8)
    m.masks = torch.ones(3, 3, 2, 100, dtype=f64)
    assert not any("/conv_f" in k for k in m.vars) and m.vars["generator/lstm/Cell0/Matrix"].shape == (200, 400)
    inp = torch.tensor(INP)
    d = m.update(inp, "d", apply=False)
    # d_loss reaches the discriminator's filters through D(fake), D(real) AND through the generator's encoder:
    m.set_requires_grad(set(m.d_vars))
    X, Y = m._split(inp)
    fake = [f.detach() for f in m.generator(X)]
    from oracle.models import DCGAN
    loss = DCGAN._ce(m.discriminator(fake), 0.0) + DCGAN._ce(m.discriminator(Y), 1.0)
    loss.backward()
    detached = m.vars["discriminator/d_conv_f2"].grad.clone()
    m.set_requires_grad(set())
    full = d["grads"]["discriminator/d_conv_f2"]
    assert (full - detached).norm() > 1e-3 * full.norm()
    # the generator's optimiser never moves the shared filters
    before = m.vars["discriminator/d_conv_f1"].clone()
    g = m.update(inp, "g")
    assert torch.equal(m.vars["discriminator/d_conv_f1"], before) and set(g["grads"]) == set(m.g_vars)


def test_variants_match_golden():
    from make_golden import RECURRENT_VARIANTS, recurrent_variant_masks
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "recurrent_variants.npz"))
    for tag, kw in RECURRENT_VARIANTS.items():
        m = RecurrentDCGAN(batch_size=2, video_length=3, seed=7, dtype=f64, **kw)
        m.masks = torch.tensor(recurrent_variant_masks())
        o = m.train_step(torch.tensor(INP))
        np.testing.assert_allclose([o["d_loss"], o["g_loss"]], g[tag + "/losses"], rtol=1e-9)
        np.testing.assert_allclose(m.vars["generator/lstm/Cell1/Bias"].numpy(), g[tag + "/final/generator/lstm/Cell1/Bias"], rtol=1e-6, atol=1e-10)
        np.testing.assert_allclose(m.vars["discriminator/d_fc_bias"].numpy(), g[tag + "/final/discriminator/d_fc_bias"], rtol=1e-6, atol=1e-10)
