"""Host-side logic of the training step (imgcomp_cvpr_b200/trainer.py) that needs no GPU: parameter packing, masks, layer
table, LR schedule, crops -- against the oracle / the reference's variable schema."""
import numpy as np
import torch

from imgcomp_cvpr_b200 import config, trainer, weights
from oracle import imgcomp_oracle as O
from oracle import train_oracle as T


def test_pc_masks_are_the_reference_masks():
    first, other = trainer._pc_masks()
    rf, ro = O.pc_masks(3)                       # create_first_mask / create_other_mask (code/probclass.py:150-176)
    assert np.array_equal(first, rf[..., 0, 0]) and np.array_equal(other, ro[..., 0, 0])
    assert first.sum() == 13 and other.sum() == 14


def test_layer_table_covers_every_conv_variable():
    for name in ('cvpr/low', 'cvpr/hi'):
        a, p = config.ae_config(name), config.pc_config('cvpr/res_shallow')
        W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
        table = trainer.ae_layer_table(a.num_chan_bn, a.arch_param_B)
        scopes = {k[:-len('/weights')] for k in W if k.startswith('autoencoder/') and k.endswith('/weights')}
        assert {t[0] for t in table} == scopes
        for s, k, stride, ci, co, tr in table:
            assert W[s + '/weights'].shape == ((k, k, co, ci) if tr else (k, k, ci, co)), s
            assert W[s + '/BatchNorm/gamma'].shape == (co,)
        assert len(table) == 2 * (6 * a.arch_param_B + 2) + 6
        for (s, _), ci, co in zip(trainer.PC_LAYERS, (1, 24, 24, 24), (24, 24, 24, a.num_centers)):
            assert W[s + '/weights'].shape == (2, 3, 3, ci, co)


def test_flat_parameter_buffer_round_trip():
    f = trainer._Flat()
    rng = np.random.RandomState(0)
    arrs = {'a': rng.randn(3, 3, 4, 8).astype(np.float32), 'b': rng.randn(5).astype(np.float32), 'c': rng.randn(130).astype(np.float32)}
    for k, v in arrs.items():
        f.add(k, v)
    f.finish(torch.device('cpu'))
    assert f.w.numel() % 64 == 0 and f.g.shape == f.w.shape == f.m.shape == f.v.shape
    for k, v in arrs.items():
        view = f.view(f.w, k)
        assert tuple(view.shape) == v.shape and np.array_equal(view.numpy(), v)
        assert view.data_ptr() % 256 == f.w.data_ptr() % 256            # 256-byte aligned relative to the buffer
    f.view(f.g, 'b')[:] = 1.0                                            # views alias the flat buffers
    assert float(f.g.sum()) == 5.0


def test_learning_rate_schedule_matches_oracle():
    a, p = config.ae_config('cvpr/med'), config.pc_config('cvpr/res_shallow')
    for cfg in (a, p):
        for step in (0, 1, 1999, 2000, 2001, 4000, 12345):
            assert trainer.learning_rate_at(cfg, step, 1000) == T.learning_rate(cfg, step, 1000)
    assert trainer.learning_rate_at(a, 1999, 1000) == a.lr_initial
    assert abs(trainer.learning_rate_at(a, 2000, 1000) - a.lr_initial * 0.1) < 1e-12
    fixed = config.Config(**dict(a.__dict__, lr_schedule='FIXED'))
    assert trainer.learning_rate_at(fixed, 10 ** 6, 1000) == a.lr_initial


def test_random_crops():
    rng = np.random.RandomState(1)
    imgs = [rng.randint(0, 256, size=(50 + 7 * i, 64 + 3 * i, 3)).astype(np.uint8) for i in range(3)]
    out = trainer.random_crops(imgs, 16, (32, 40), np.random.RandomState(2))
    assert out.shape == (16, 3, 32, 40) and out.dtype == np.uint8
    # every crop (or its horizontal flip) occurs in one of the images
    for c in out[:4]:
        hwc = c.transpose(1, 2, 0)
        found = False
        for im in imgs:
            for cand in (hwc, hwc[:, ::-1]):
                H, W = im.shape[:2]
                for y in range(H - 32 + 1):
                    rows = np.where((im[y, :W - 40 + 1, 0] == cand[0, 0, 0]))[0]
                    for x in rows:
                        if np.array_equal(im[y:y + 32, x:x + 40], cand):
                            found = True
        assert found


def test_tape_accumulation_rules():
    """gradients that arrive twice are summed; tensors the tape does not own are never modified in place"""
    calls = []

    class FakeNN(object):
        @staticmethod
        def axpby(a, x, b=0.0, y=None, out=None):
            r = a * x + (b * y if y is not None else 0)
            if out is not None:
                out.copy_(r)
                return out
            return r

        @staticmethod
        def add(x, y):
            calls.append('add')
            return x + y
    real = trainer.nn
    trainer.nn = FakeNN
    try:
        tape = trainer._Tape()
        t = torch.zeros(3)
        g1, g2 = torch.ones(3), torch.full((3,), 2.0)
        tape.acc(t, g1)                  # not owned
        tape.acc(t, g2)                  # not owned either -> a fresh tensor
        assert calls == ['add'] and float(g1.sum()) == 3.0 and float(g2.sum()) == 6.0
        g3 = torch.full((3,), 4.0)
        tape.acc(t, g3, owned=True)      # accumulates into the tape's own tensor
        assert float(tape.pop(t).sum()) == 21.0 and tape.pop(t) is None
        order = []
        tape.add(lambda: order.append(1))
        tape.add(lambda: order.append(2))
        tape.backward()
        assert order == [2, 1]
    finally:
        trainer.nn = real


def _cpu_trainer(ae_name='cvpr/low'):
    a, p = config.ae_config(ae_name), config.pc_config('cvpr/res_shallow')
    W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    return a, p, W, trainer.Trainer(a, p, W, num_itr_per_epoch=100, device='cpu')


def test_parameter_packing_round_trip():
    """TF-named variables -> padded, op-oriented flat device buffers -> back: exact, for every variable (transposed convs
    swap their channel axes on the way in and out, 3 / C+1 / L channel counts are padded to multiples of 4)"""
    for name in ('cvpr/low', 'cvpr/hi'):
        a, p, W, tr = _cpu_trainer(name)
        out = tr.weights()
        assert sorted(out) == sorted(W)
        for k in W:
            assert out[k].shape == np.asarray(W[k]).shape and np.array_equal(out[k], W[k]), k
        # the context-model masks sit next to the padded weights and match the reference masks
        first, other = trainer._pc_masks()
        m0 = tr._const('probclass3d/logits/conv3d_conv0_mask/mask').numpy()
        assert m0.shape == (2, 3, 3, 4, 24) and np.array_equal(m0[..., 0, 0], first) and m0[..., 1:, :].sum() == 0
        m3 = tr._const('probclass3d/logits/conv3d_conv2_mask/mask').numpy()
        assert m3.shape == (2, 3, 3, 24, 8) and np.array_equal(m3[..., 0, 0], other) and m3[..., 6:].sum() == 0
        # padded entries of the device weights are zero (so their gradients and Adam updates stay zero)
        w_h1 = tr._w('autoencoder/encoder/h1/weights')
        assert tuple(w_h1.shape) == (5, 5, 4, 64) and float(w_h1[:, :, 3].abs().sum()) == 0.0
        w_h13 = tr._w('autoencoder/decoder/h13/weights')
        assert tuple(w_h13.shape) == (5, 5, 64, 4) and float(w_h13[..., 3].abs().sum()) == 0.0


def test_state_dict_round_trip():
    a, p, W, tr = _cpu_trainer()
    rng = np.random.RandomState(3)
    for g in tr.groups.values():             # pretend some steps happened
        g.m.copy_(torch.from_numpy(rng.randn(g.m.numel()).astype(np.float32)))
        g.v.copy_(torch.from_numpy(np.abs(rng.randn(g.v.numel())).astype(np.float32)))
    tr.global_step, tr._adam_t = 1234, 1234
    sd = tr.state_dict()
    assert 'autoencoder/encoder/h2/weights/Adam' in sd and 'probclass3d/logits/conv3d_conv0_mask/biases/Adam_1' in sd
    _, _, _, tr2 = _cpu_trainer()
    tr2.load_state_dict(sd)
    assert tr2.global_step == 1234 and tr2._adam_t == 1234
    sd2 = tr2.state_dict()
    assert sorted(sd) == sorted(sd2)
    for k in sd:
        assert np.array_equal(sd[k], sd2[k]), k
    # partial restore (restore_skip_vars-like): only the centres
    _, _, _, tr3 = _cpu_trainer()
    tr3.load_state_dict({'autoencoder/encoder/centers': np.arange(6, dtype=np.float32)})
    assert np.array_equal(tr3.weights()['autoencoder/encoder/centers'], np.arange(6, dtype=np.float32))
    assert np.array_equal(tr3.weights()['autoencoder/encoder/h1/weights'], W['autoencoder/encoder/h1/weights'])
