"""Counter-based source of the shared uniform / normal buffers (SURVEY.md App. A
"shared-uniform-buffer mode").  TEST INFRASTRUCTURE ONLY.

The reference draws its randomness from Theano's MRG31k3p streams, which this
build does not reproduce; parity is defined with both sides reading the SAME
buffer.  A buffer is a pure function of (seed, layer, call index), so the
reference-under-shim, the oracle and the CUDA path can each ask for "the
randomness of step t of layer l" without sharing state.

Uniforms are (n + 0.5) / 2**24 with n in [0, 2**24): exactly representable in
fp32, strictly inside (0, 1), so `u < p` never depends on rounding of u."""
import numpy as np

from . import rbm_oracle as O


def _gen(seed, layer, call):
    return np.random.Generator(np.random.Philox(key=np.random.SeedSequence(
        [int(seed), int(layer), int(call)]).generate_state(2, dtype=np.uint64)))


def uniforms(g, n):
    return ((g.integers(0, 1 << 24, size=n, dtype=np.int64).astype(np.float64) + 0.5)
            / float(1 << 24)).astype(np.float32)


def step_buffer(seed, layer, call, kind, error_free, B, V, H, k):
    """Flat fp32 buffer in the App. A layout; N_v segments (noisy GRBM) hold
    standard normals, every other segment uniforms."""
    g = _gen(seed, layer, call)
    lay, n = O.u_layout(kind, error_free, B, V, H, k)
    buf = np.empty(n, np.float32)
    for name, off, sh in lay:
        m = sh[0] * sh[1]
        if kind == O.GRBM and name.startswith("v"):
            buf[off:off + m] = g.standard_normal(m).astype(np.float32)
        else:
            buf[off:off + m] = uniforms(g, m)
    return buf
