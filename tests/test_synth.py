import numpy as np
import torch

from gemma_b200 import synth
from oracle import oracle as O


def test_numpy_and_torch_generators_are_bit_identical():
    for n, l, miss in ((37, 9, 0.0), (130, 17, 0.05)):
        bed, G = synth.make_bed(n, l, seed=5, snp_offset=3, miss_rate=miss)
        bt = synth.make_bed_torch(n, l, "cpu", seed=5, snp_offset=3, miss_rate=miss, chunk=4)
        assert np.array_equal(bed, bt.numpy())
        for s in range(l):
            g = O.bed_decode(bed[s].tobytes(), n)
            ref = np.where(G[s] < 0, np.nan, G[s])
            assert np.array_equal(np.isnan(g), np.isnan(ref)) and np.array_equal(np.nan_to_num(g), np.nan_to_num(ref))


def test_allele_frequencies_in_range():
    g = synth.genotypes(4000, 50, seed=1)
    f = g.mean(axis=1) / 2
    assert f.min() > 0.02 and f.max() < 0.55 and set(np.unique(g)) <= {0, 1, 2}
