"""Executable statement of the shared-memory operand layouts the tcgen05 kernels rely on (csrc/conv_tc.cu,
csrc/train_tc.cu), checked on the CPU with a small emulator of the UMMA SWIZZLE_NONE shared-memory descriptor
(cute/atom/mma_traits_sm100.hpp "canonical layouts", in 16-byte units):

    K-major  ((8,n),2):((1,SBO),LBO)        element (r, k) at  start + (r//8) SBO + (k//8) LBO + (r%8) 16 + (k%8) 2
    MN-major ((1,n),(8,k)):((X,SBO),(1,LBO)) element (r, k) at  start + (r//8) SBO + (k//8) LBO + (k%8) 16 + (r%8) 2

(r = M or N index, k = K index, fp16).  The tests rebuild the tiles exactly as TMA delivers them from the NC/8HW8
planes, form the descriptors the way the kernels do, and compare the emulated operand matrices / products with plain
numpy: the 3x3 conv (tap = 16-byte shift of A's start address), its weights stage, the filter gradient as a GEMM over
pixels on MN-major operands, and the B-concatenation layout planned in DESIGN.md section 7 (so the next kernel can be
desk-checked before it sees a GPU)."""
import numpy as np
import pytest


def operand(smem, start, lbo, sbo, rows, major, kdim=16):
    """the (rows x kdim) fp16 matrix one MMA reads through a SWIZZLE_NONE descriptor (smem: uint8 array)"""
    assert start % 16 == 0 and lbo % 16 == 0 and sbo % 16 == 0, 'descriptor fields are in 16-byte units'
    h = smem.view(np.float16)
    out = np.empty((rows, kdim), np.float16)
    for r in range(rows):
        for k in range(kdim):
            if major == 'K':
                a = start + (r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2
            else:
                a = start + (r // 8) * sbo + (k // 8) * lbo + (k % 8) * 16 + (r % 8) * 2
            out[r, k] = h[a // 2]
    return out


def tma_tile(planes, n, chunk0, nchunks, y0, x0, th, tw):
    """what a TMA box {tw*8, th, nchunks, 1, 1} at (x0, y0, chunk0, n) of a [N][C/8][H][W][8] plane delivers:
    [chunk][row][pixel][8] fp16, out-of-bounds elements zero"""
    N, CH, H, W, _ = planes.shape
    out = np.zeros((nchunks, th, tw, 8), np.float16)
    for c in range(nchunks):
        for r in range(th):
            for p in range(tw):
                y, x, ch = y0 + r, x0 + p, chunk0 + c
                if 0 <= y < H and 0 <= x < W and 0 <= ch < CH:
                    out[c, r, p] = planes[n, ch, y, x]
    return out


def to_planes(x_nhwc):
    """float NHWC -> fp16 [N][C/8][H][W][8] (the hi plane; values chosen exactly representable)"""
    N, H, W, C = x_nhwc.shape
    return np.ascontiguousarray(x_nhwc.reshape(N, H, W, C // 8, 8).transpose(0, 3, 1, 2, 4).astype(np.float16))


def _ints(rng, shape, lo=-4, hi=5):
    return rng.randint(lo, hi, size=shape).astype(np.float32)


TH, TW, T = 16, 8, 2                                    # conv_tc.cu: one MMA tile = 16 rows x 8 pixels, T tiles per super tile
HALO_W, HALO_H = TW * T + 2, TH + 2
HALO_PIX = HALO_W * HALO_H


def test_conv3x3_tile_taps_are_descriptor_shifts():
    rng = np.random.RandomState(0)
    N, H, W, Cin, Cout = 1, 20, 19, 32, 128             # ragged: the tile hangs over the right / bottom edge
    x = _ints(rng, (N, H, W, Cin))
    w = _ints(rng, (3, 3, Cin, Cout), -2, 3)
    planes = to_planes(x)
    y0, x0 = 0, 0
    a_tile = tma_tile(planes, 0, 0, 4, y0 - 1, x0 - 1, HALO_H, HALO_W)          # halo0 = -1: SAME padding by OOB zero fill
    a_smem = a_tile.reshape(4, HALO_PIX, 8).view(np.uint8).reshape(-1)          # [chunk][halo pixel][8]
    # one weight stage per tap: [4 chunks][Cout rows][8 cin]
    acc = np.zeros((T, 128, Cout), np.float64)
    for tap in range(9):
        dy, dx = divmod(tap, 3)
        w_stage = np.zeros((4, Cout, 8), np.float16)
        for ch in range(4):
            w_stage[ch] = w[dy, dx, ch * 8:(ch + 1) * 8, :].T
        w_smem = w_stage.view(np.uint8).reshape(-1)
        for t in range(T):
            for ks in range(2):
                a = operand(a_smem, (dy * HALO_W + dx + t * TW) * 16 + ks * 2 * HALO_PIX * 16, lbo=HALO_PIX * 16, sbo=HALO_W * 16,
                            rows=128, major='K')
                b = operand(w_smem, ks * 2 * Cout * 16, lbo=Cout * 16, sbo=128, rows=Cout, major='K')
                acc[t] += a.astype(np.float64) @ b.astype(np.float64).T
    # reference: SAME 3x3 conv at the tile's pixels (row m of the tile = pixel (m >> 3, (m & 7) + 8 t))
    xp = np.pad(x[0], ((1, 1 + TH), (1, 1 + TW * T), (0, 0)))
    for t in range(T):
        for m in range(128):
            ty, tx = m >> 3, (m & 7) + TW * t
            ref = sum(xp[y0 + ty + dy, x0 + tx + dx] @ w[dy, dx] for dy in range(3) for dx in range(3))
            if y0 + ty < H and x0 + tx < W:
                assert np.array_equal(acc[t, m], ref), (t, m)


WG_ROWS, WG_COLS = 4, 16                                # train_tc.cu: filter-gradient pixel tile


def test_filter_gradient_is_a_gemm_over_pixels_on_mn_major_tiles():
    rng = np.random.RandomState(1)
    N, H, W, C = 2, 6, 21, 128                          # H not a multiple of 4, W not of 16
    x = _ints(rng, (N, H, W, C), -3, 4)
    dy = _ints(rng, (N, H, W, C), -3, 4)
    xp, dyp = to_planes(x), to_planes(dy)
    dw = np.zeros((3, 3, C, C), np.float64)
    for ky in range(3):                                 # CTA (ky, split): taps (ky, 0..2); here one "split" walks every tile
        for n in range(N):
            for y0 in range(0, H, WG_ROWS):
                for x0 in range(0, W, WG_COLS):
                    xt = tma_tile(xp, n, 0, 16, y0 + ky - 1, x0 - 1, WG_ROWS, WG_COLS + 2).view(np.uint8).reshape(-1)
                    yt = tma_tile(dyp, n, 0, 16, y0, x0, WG_ROWS, WG_COLS).view(np.uint8).reshape(-1)
                    for r in range(WG_ROWS):
                        b = operand(yt, r * WG_COLS * 16, lbo=128, sbo=WG_ROWS * WG_COLS * 16, rows=C, major='MN')       # (cout, pixel)
                        for kx in range(3):
                            a = operand(xt, (r * (WG_COLS + 2) + kx) * 16, lbo=128, sbo=WG_ROWS * (WG_COLS + 2) * 16, rows=C,
                                        major='MN')                                                                        # (cin, pixel)
                            dw[ky, kx] += a.astype(np.float64) @ b.astype(np.float64).T
    # reference: dW[ky][kx][ci][co] = sum_p x[p + (ky-1, kx-1)][ci] dy[p][co]
    xpad = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0))).astype(np.float64)
    for ky in range(3):
        for kx in range(3):
            ref = np.einsum('nyxi,nyxo->io', xpad[:, ky:ky + H, kx:kx + W], dy.astype(np.float64))
            assert np.array_equal(dw[ky, kx], ref), (ky, kx)


def test_b_concatenation_layout_for_exact_mode():
    """weights of one stage stored [chunk][plane hi|lo][rows][8] make [w_hi | w_lo] ONE K-major B operand
    of N = 2 * rows (LBO = 2 * rows * 16), and w_hi alone the first `rows` rows of the same descriptor."""
    rng = np.random.RandomState(2)
    rows = 128
    w_hi, w_lo = _ints(rng, (32, rows)), _ints(rng, (32, rows))                      # (cin within the group, cout)
    stage = np.zeros((4, 2, rows, 8), np.float16)
    for ch in range(4):
        stage[ch, 0] = w_hi[ch * 8:(ch + 1) * 8].T
        stage[ch, 1] = w_lo[ch * 8:(ch + 1) * 8].T
    smem = stage.view(np.uint8).reshape(-1)
    for ks in range(2):
        b = operand(smem, ks * 2 * (2 * rows * 16), lbo=2 * rows * 16, sbo=128, rows=2 * rows, major='K')
        k = slice(ks * 16, ks * 16 + 16)
        assert np.array_equal(b[:rows], w_hi[k].T) and np.array_equal(b[rows:], w_lo[k].T)
        assert np.array_equal(operand(smem, ks * 2 * (2 * rows * 16), lbo=2 * rows * 16, sbo=128, rows=rows, major='K'), w_hi[k].T)


def test_context_model_third_chunks_of_two_taps_share_one_k_step():
    """conv_tc.cu, context-model kernels (pair_c2): 24 input channels are 3 chunks, so k-step 1 of a tap would hold one real
    chunk and one all-zero chunk.  The third chunks of taps 2j and 2j+1 form ONE k-step instead: A's two K core matrices are
    (tap 2j, chunk 2) and (tap 2j+1, chunk 2) -- LBO = the tap shift, 16 bytes for horizontally adjacent taps, i.e. overlapping
    core matrices -- and B's are chunk 2 of the two taps' resident stages, LBO = the stage pitch.  The emulated schedule
    (22 k-steps for the 14 taps) must equal the VALID masked conv."""
    rng = np.random.RandomState(3)
    Hh, Ww, K, NO = 18, 10, 24, 32                      # one 16 x 8 tile with its VALID halo; 24 channels, 32 columns
    x = _ints(rng, (2, Hh, Ww, 32))                     # two depth slices, channels padded to 32
    x[..., K:] = 0
    w = _ints(rng, (2, 3, 3, 32, NO), -2, 3)
    w[:, :, :, K:, :] = 0
    halo_w, halo_pix = Ww, Hh * Ww
    groups = [[(0, fy, fx) for fy in range(3) for fx in range(3)],
              [(1, fy, fx) for fy in range(3) for fx in range(3) if not (fy > 1 or (fy == 1 and fx > 1))]]
    acc = np.zeros((128, NO), np.float64)
    nks = 0
    for g, taps in enumerate(groups):
        a_smem = tma_tile(to_planes(x[g:g + 1]), 0, 0, 4, 0, 0, Hh, Ww).reshape(4, halo_pix, 8).view(np.uint8).reshape(-1)
        # resident weight stages of the group, one per tap: [4 chunks][NO rows][8 cin]
        stages = np.zeros((len(taps), 4, NO, 8), np.float16)
        for ti, (fd, fy, fx) in enumerate(taps):
            for ch in range(4):
                stages[ti, ch] = w[fd, fy, fx, ch * 8:(ch + 1) * 8, :].T
        w_smem = stages.view(np.uint8).reshape(-1)
        stage_pitch, chunk_pitch = 4 * NO * 16, NO * 16
        shift = lambda t: (t[1] * halo_w + t[2]) * 16
        for ti, tap in enumerate(taps):
            if ti & 1:                                   # chunk 2 of taps ti-1 and ti
                prev = taps[ti - 1]
                a = operand(a_smem, shift(prev) + 2 * halo_pix * 16, lbo=shift(tap) - shift(prev), sbo=halo_w * 16, rows=128, major='K')
                b = operand(w_smem, (ti - 1) * stage_pitch + 2 * chunk_pitch, lbo=stage_pitch, sbo=128, rows=NO, major='K')
                acc += a.astype(np.float64) @ b.astype(np.float64).T
                nks += 1
            a = operand(a_smem, shift(tap), lbo=halo_pix * 16, sbo=halo_w * 16, rows=128, major='K')      # chunks 0, 1
            b = operand(w_smem, ti * stage_pitch, lbo=chunk_pitch, sbo=128, rows=NO, major='K')
            acc += a.astype(np.float64) @ b.astype(np.float64).T
            nks += 1
            if (ti & 1) == 0 and ti == len(taps) - 1:    # unpaired last tap: chunk 2 with the all-zero chunk 3
                a = operand(a_smem, shift(tap) + 2 * halo_pix * 16, lbo=halo_pix * 16, sbo=halo_w * 16, rows=128, major='K')
                b = operand(w_smem, ti * stage_pitch + 2 * chunk_pitch, lbo=chunk_pitch, sbo=128, rows=NO, major='K')
                acc += a.astype(np.float64) @ b.astype(np.float64).T
                nks += 1
    assert nks == 22
    for m in range(128):
        ty, tx = m >> 3, m & 7
        ref = sum(x[fd, ty + fy, tx + fx].astype(np.float64) @ w[fd, fy, fx] for taps in groups for (fd, fy, fx) in taps)
        assert np.array_equal(acc[m], ref), m


def _fast_operand(smem, start, lbo, sbo, rows):
    """operand(..., major='K') vectorised (the walk test issues a few hundred MMAs)"""
    h = smem.view(np.float16)
    r = np.arange(rows)[:, None]
    k = np.arange(16)[None, :]
    a = start + (r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2
    return h[a // 2]


@pytest.mark.parametrize('NO,pairing', [(32, True), (16, True), (32, False)])
def test_context_model_depth_walk_schedule(NO, pairing):
    """conv_tc.cu, ConvTcParams::walk: a CTA follows one (image, tile) along the depth axis; input slice j FINISHES output
    j - 1 (filter depth 1, taps 0..4) and STARTS output j (filter depth 0: taps 5..8, then 0..4) from ONE activation tile.
    Accumulators: ring of 4 tiles [X | Y] of 2 NO TMEM columns, output = X + Y.  Taps 0..4 are stored
    [4 chunks][W1lo | W1hi | W0hi | W0lo rows][8] (built from the global per-tap stages [4 chunks][hi | lo][8] by the same
    copies as the kernel's weight producer), so a_hi x all 4 NO rows lands on [X_f | Y_f | X_s | Y_s] and a_lo x the middle
    2 NO rows on [Y_f | X_s].  The emulation issues exactly the kernel's descriptors (tap shifts, paired third chunks, row
    offsets, separate halves where the ring wraps and at segment ends) and must reproduce
    a_hi w_hi + a_hi w_lo + a_lo w_hi of the VALID masked 3-D conv for every output slice."""
    rng = np.random.RandomState(5)
    K, Din = 24, 8                                      # 7 output slices: the ring of 4 wraps once; NO = 32: layers 1-2, 16: the head
    Hh, Ww = 18, 10
    halo_w, halo_pix = Ww, Hh * Ww
    a_plane = 4 * halo_pix * 16
    xh = _ints(rng, (Din, Hh, Ww, 32)); xl = _ints(rng, (Din, Hh, Ww, 32), -2, 3)
    xh[..., K:] = 0; xl[..., K:] = 0
    wh = _ints(rng, (2, 3, 3, 32, NO), -2, 3); wl = _ints(rng, (2, 3, 3, 32, NO), -1, 2)
    wh[:, :, :, K:, :] = 0; wl[:, :, :, K:, :] = 0
    taps_fd = [[(fy, fx) for fy in range(3) for fx in range(3)], [(fy, fx) for fy in range(3) for fx in range(3)][:5]]
    # global stages as pack_weights_pc writes them: 9 stages of filter depth 0, 5 of depth 1; [4 chunks][hi rows | lo rows][8]
    S1 = 4 * 2 * NO * 16
    gl = np.zeros((14, 4, 2 * NO, 8), np.float16)
    for fd in range(2):
        for ti, (fy, fx) in enumerate(taps_fd[fd]):
            for ch in range(4):
                gl[fd * 9 + ti, ch, :NO] = wh[fd, fy, fx, ch * 8:(ch + 1) * 8, :].T
                gl[fd * 9 + ti, ch, NO:] = wl[fd, fy, fx, ch * 8:(ch + 1) * 8, :].T
    g8 = gl.view(np.uint8).reshape(14, S1)
    # the weight producer's bulk copies
    w_smem = np.zeros(14 * S1, np.uint8)
    rows = NO * 16
    for t in range(5):
        for c in range(4):
            dst = t * 2 * S1 + c * 4 * rows
            w0, w1 = g8[t, c * 2 * rows:], g8[9 + t, c * 2 * rows:]
            w_smem[dst:dst + rows] = w1[rows:2 * rows]
            w_smem[dst + rows:dst + 2 * rows] = w1[:rows]
            w_smem[dst + 2 * rows:dst + 4 * rows] = w0[:2 * rows]
    for t in range(5, 9):
        w_smem[10 * S1 + (t - 5) * S1:10 * S1 + (t - 4) * S1] = g8[t]
    tmem = np.zeros((128, 8 * NO), np.float64)
    nmma = [0]

    def umma(d, n, a_start, a_lbo, b_start, b_lbo, acc, a_smem):
        a = _fast_operand(a_smem, a_start, a_lbo, halo_w * 16, 128).astype(np.float64)
        b = _fast_operand(w_smem, b_start, b_lbo, 128, n).astype(np.float64)
        assert d + n <= tmem.shape[1]
        tmem[:, d:d + n] = (tmem[:, d:d + n] if acc else 0) + a @ b.T
        nmma[0] += 1

    def issue_taps(nt, tap0, cat2, a_smem, w_g, d_hi, n_hi, d_lo, w_lo_off, n_lo, first):
        stage = 2 * S1 if cat2 else S1
        wlbo = (4 if cat2 else 2) * NO * 16
        wks, aks, albo = 2 * wlbo, 2 * halo_pix * 16, halo_pix * 16
        for ti in range(nt):
            tap = tap0 + ti
            a_t = ((tap // 3) * halo_w + tap % 3) * 16
            w_t = w_g + ti * stage
            if pairing and (ti & 1):
                a_prev = (((tap - 1) // 3) * halo_w + (tap - 1) % 3) * 16
                umma(d_hi, n_hi, a_prev + aks, a_t - a_prev, w_t - stage + wks, stage, True, a_smem)
                umma(d_lo, n_lo, a_prev + aks + a_plane, a_t - a_prev, w_t - stage + wks + w_lo_off * 16, stage, True, a_smem)
            umma(d_hi, n_hi, a_t, albo, w_t, wlbo, not (first and ti == 0), a_smem)
            umma(d_lo, n_lo, a_t + a_plane, albo, w_t + w_lo_off * 16, wlbo, True, a_smem)
            if not pairing or ((ti & 1) == 0 and ti == nt - 1):
                umma(d_hi, n_hi, a_t + aks, albo, w_t + wks, wlbo, True, a_smem)
                umma(d_lo, n_lo, a_t + aks + a_plane, albo, w_t + wks + w_lo_off * 16, wlbo, True, a_smem)

    def reference(d):
        out = np.zeros((128, NO), np.float64)
        for m in range(128):
            ty, tx = m >> 3, m & 7
            for fd in range(2):
                for (fy, fx) in taps_fd[fd]:
                    ah, al = xh[d + fd, ty + fy, tx + fx].astype(np.float64), xl[d + fd, ty + fy, tx + fx].astype(np.float64)
                    out[m] += ah @ wh[fd, fy, fx] + ah @ wl[fd, fy, fx] + al @ wh[fd, fy, fx]
        return out

    w2, w1 = 0, 10 * S1
    results = {}
    for seg_len in (7, 3):                               # one segment per column; three segments (3 + 3 + 1 outputs)
        tmem[:] = np.nan                                 # a start must overwrite, never accumulate onto stale columns
        nacc, done = 0, []
        for da in range(0, Din - 1, seg_len):
            db = min(Din - 1, da + seg_len)
            for j in range(da, db + 1):
                start, finish = j < db, j > da
                ss, sf = nacc & 3, (nacc - 1) & 3
                tile = np.concatenate([tma_tile(to_planes(v[j:j + 1]), 0, 0, 4, 0, 0, Hh, Ww).reshape(-1) for v in (xh, xl)])
                a_smem = tile.view(np.uint8)
                d_s, d_f = ss * 2 * NO, sf * 2 * NO
                if start:
                    issue_taps(4, 5, False, a_smem, w1, d_s, 2 * NO, d_s, 0, NO, True)
                if start and finish and ss != 0:
                    issue_taps(5, 0, True, a_smem, w2, d_f, 4 * NO, d_f + NO, NO, 2 * NO, False)
                else:
                    if start:
                        issue_taps(5, 0, True, a_smem, w2 + 2 * NO * 16, d_s, 2 * NO, d_s, 0, NO, False)
                    if finish:
                        issue_taps(5, 0, True, a_smem, w2, d_f, 2 * NO, d_f + NO, NO, NO, False)
                if finish:
                    done.append((j - 1, (tmem[:, d_f:d_f + NO] + tmem[:, d_f + NO:d_f + 2 * NO]).copy()))
                if start:
                    nacc += 1
        assert [d for d, _ in done] == list(range(Din - 1))
        for d, got in done:
            assert np.array_equal(got, reference(d)), (seg_len, d)
        results[seg_len] = nmma[0]
        nmma[0] = 0
    # k-steps (x 2 planes): start half 14, finish half 8, fused step 14, wrapped step 14 + 8 -- against 22 per output of the
    # two-group schedule.  One segment of 7 outputs: start + 5 fused + 1 wrapped + finish; three segments: 3 starts, 3 finishes,
    # 3 fused + 1 wrapped (the ring position carries over from segment to segment)
    if pairing:
        assert results[7] == 2 * (14 + 5 * 14 + 22 + 8), results
        assert results[3] == 2 * (3 * 14 + 3 * 8 + 3 * 14 + 22), results
        assert results[7] < 2 * 22 * 7
    else:       # IC_PC_PAIR_CHUNKS=0: two k-steps per tap, 18 per fused step, 18 + 10 where the halves are separate
        assert results[7] == 2 * (18 + 5 * 18 + 28 + 10), results
