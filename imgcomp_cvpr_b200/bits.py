"""Host mirror of code/bits.py."""


def bitcost_to_bpp(bit_cost, input_batch):
    """:param bit_cost: NChw  :param input_batch: N3HW  :return: num_bits / num_pixels
    (code/bits.py:4-14) -- one scalar for the batch, like the reference."""
    assert bit_cost.dim() == input_batch.dim() == 4, 'Expected NChw and N3HW'
    return bit_cost.sum() / float(num_pixels_in_input_batch(input_batch))


def num_pixels_in_input_batch(input_batch):
    """code/bits.py:17-20"""
    assert int(input_batch.shape[1]) == 3, 'Expected N3HW, got {}'.format(tuple(input_batch.shape))
    return input_batch.numel() / 3
