"""N>1 host logic on CPU: contiguous sharding + the single metric all-reduce, world size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imgcomp_cvpr_b200 import val


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 24, 25, 64):
        for world in (1, 2, 3, 8):
            parts = [val.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_add_padding_matches_reference_rule():
    im = np.arange(5 * 13 * 3, dtype=np.uint8).reshape(5, 13, 3)
    p, undo = val.add_padding(im, 8)
    assert p.shape == (8, 16, 3)
    # centred: (8-5)=3 -> 1 before / 2 after ; (16-13)=3 -> 1 before / 2 after   (images_iterator.py:45-55)
    assert np.array_equal(p[1:6, 1:14], im) and p[0].sum() == 0 and p[:, 0].sum() == 0
    assert np.array_equal(undo(p), im)
    q, undo2 = val.add_padding(np.zeros((16, 8, 4), np.uint8), 8)
    assert q.shape == (16, 8, 3)


def _worker(rank, world, port, n_items, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    table = rng.uniform(0.1, 40, size=(n_items, 3))
    lo, hi = val.shard_range(n_items, rank, world)
    rows = [dict(zip(val.METRICS, table[i])) for i in range(lo, hi)]
    avgs, n = val.reduce_metric_sums(rows, dist)
    if rank == 0:
        out.put((avgs, n, table.mean(0).tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_items', [5, 24])
def test_two_rank_metric_reduce(n_items):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    avgs, n, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n == n_items
    for k, w in zip(val.METRICS, want):
        assert abs(avgs[k] - w) < 1e-12


def test_nan_metric_is_rejected():
    with pytest.raises(AssertionError):
        val.reduce_metric_sums([{'bpp': float('nan'), 'ms-ssim': 1.0, 'psnr': 30.0}])


def test_measures_writer_format(tmp_path):
    """measures.csv as the reference's MeasuresWriter / MeasuresReader exchange it (code/val_files.py:62-100)"""
    from imgcomp_cvpr_b200 import val
    w = val.MeasuresWriter(str(tmp_path))
    w.append('kodim01', {'bpp': 0.25, 'ms-ssim': 0.97, 'psnr': 30.5})
    w.close()
    lines = open(str(tmp_path / 'measures.csv')).read().splitlines()
    assert lines[0] == 'img_name,bpp,ms-ssim,psnr'
    name, bpp, ms, ps = lines[1].split(',')
    assert name == 'kodim01' and float(bpp) == 0.25 and float(ms) == 0.97 and float(ps) == 30.5


def test_save_img_writes_the_output_png(tmp_path):
    """val.save_img (code/val.py:215-225): CHW uint8 -> <out_dir>/imgs/<name>.png, read back identical"""
    from PIL import Image
    from imgcomp_cvpr_b200 import val
    img = np.random.RandomState(0).randint(0, 256, size=(3, 24, 40)).astype(np.uint8)
    for name in ('kodim01', 'kodim02.png'):
        p = val.save_img(name, img, str(tmp_path))
        assert p.endswith('.png') and os.path.dirname(p) == os.path.join(str(tmp_path), 'imgs')
        back = np.asarray(Image.open(p))
        assert np.array_equal(back, img.transpose(1, 2, 0))
