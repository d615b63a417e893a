"""Real (arithmetic-coded) vs theoretical bits per pixel of one image's symbols: the glue val.py uses for
--real_bpp (reference: code/bpp_helpers.py:5-31, driven from code/val.py:161-175).

Both numbers come from ONE batched context-model pass here: `PredictionNetwork.get_all_freqs` returns every
position's frequency table together with the cross-entropy bit cost of the same pass, the host range coder
then produces the actual stream length."""
from . import bit_counter


class BppFetcher(object):
    def __init__(self, pred, checker):
        """pred: probclass.PredictionNetwork, checker: probclass.ProbclassNetworkTesting (same roles as in the
        reference constructor, code/bpp_helpers.py:9-11)."""
        self.pred, self.checker = pred, checker

    def get_bpp(self, symbols, num_pixels):
        """symbols: ndarray N x C x h x w of one image (N = 1 in val.py) -> (bpp_real, bpp_theory):
        coded stream bits / pixels and sum(-log2 p) / pixels (code/bpp_helpers.py:13-25)."""
        if symbols.ndim != 4:
            raise AssertionError('expected NCHW symbols, got shape {}'.format(symbols.shape))
        coded_bits = 0
        for per_image in symbols:                       # batch entries are coded independently and summed
            coded_bits += bit_counter.encode_decode_to_file_ctx(per_image, self.pred, syms_format='CHW', verbose=True)
        theoretical_bits = self.checker.get_total_bit_cost(symbols)
        return coded_bits / num_pixels, theoretical_bits / num_pixels


def num_pixels_in_image(im):
    """im: CHW RGB array -> H * W (code/bpp_helpers.py:28-31)."""
    if im.shape[0] != 3:
        raise AssertionError('Expected RGB image, got {}'.format(im.shape))
    return im.shape[1] * im.shape[2]
