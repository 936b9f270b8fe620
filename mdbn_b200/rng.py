"""Randomness sources standing in for theano_rng (src/rbm.py:40,92; src/dbn.py:114).

RandomStreams(seed)      production: Philox4x32-10 evaluated inside the kernels.
BufferStreams(provider)  parity mode: every CD step reads a caller-supplied
                         flat fp32 buffer of uniforms / normals, the same one the CPU oracle reads."""
import numpy as np
import torch

from . import _lib


class RandomStreams:
    mode = _lib.RNG_PHILOX

    def __init__(self, seed=None):
        self.seed = int(seed) if seed is not None else 12345
        self._site_offsets = {}
        self._site_ids = {}

    def next_rng(self, site, device, n_values=0, n_steps=1, **_):
        """mdbn_rng for the next call (or the next `n_steps` chained calls) of sampling site `site`.
        Sites are numbered in order of first use, so a program that makes the same calls in the same order
        draws the same numbers in every run (the object id itself never enters the stream)."""
        sid = self._site_ids.setdefault(site, len(self._site_ids))
        off = self._site_offsets.get(site, 0)
        self._site_offsets[site] = off + int(n_steps)
        # decorrelate sites by folding the site number into the upper half of the key
        return _lib.Rng(_lib.RNG_PHILOX, None, (self.seed ^ ((sid * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF) << 32) & (2 ** 64 - 1), off), None


class BufferStreams:
    """provider(layer_id, call_idx, B) -> 1-D float32 array laid out per App. A."""
    mode = _lib.RNG_BUFFER

    def __init__(self, provider):
        self.provider = provider
        self._calls = {}

    def next_rng(self, site, device, layer_id=0, B=0, **_):
        n = self._calls.get(site, 0)
        self._calls[site] = n + 1
        buf = np.ascontiguousarray(self.provider(layer_id, n, B), dtype=np.float32)
        t = torch.from_numpy(buf).to(device)
        return _lib.Rng(_lib.RNG_BUFFER, t.data_ptr(), 0, 0), t   # keep t alive until the step is enqueued


class ExplicitBuffer:
    """One-shot buffer for the single-phase calls (sample_h_given_v(v, u=...))."""
    mode = _lib.RNG_BUFFER

    def __init__(self, values):
        self.values = values

    def next_rng(self, site, device, **_):
        t = torch.as_tensor(np.ascontiguousarray(self.values, dtype=np.float32)).to(device)
        return _lib.Rng(_lib.RNG_BUFFER, t.data_ptr(), 0, 0), t
