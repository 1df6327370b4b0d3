"""JAX-compatible counter-based PRNG on the host.

The reference derives every random stream from `jax.random` keys (threefry2x32, the default
`jax_default_prng_impl` of the pinned jax==0.4.26, with `jax_threefry_partitionable` off):
`split` / `fold_in` for the per-member keys (inference.py:436-441, :571-575, :618), and
`permutation` for the per-epoch batch order (`permute_dataset`, inference.py:35-39, :591-595).
JAX is not installable in this environment, so this module restates the published algorithms in
numpy:

* Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11;
  rotation constants and key schedule as in Random123 / `jax._src.prng.threefry2x32_p`),
* `jax._src.prng`: `threefry_seed`, `threefry_split`, `threefry_fold_in`, `threefry_random_bits`
  (original, non-partitionable layout: counts = iota(n), odd n padded with one zero, the two
  halves of the count vector are the two cipher words, outputs concatenated),
* `jax._src.random`: `_uniform` (mantissa trick), `_normal_real` (sqrt(2)*erfinv(u)), `_shuffle`
  (ceil(3 ln n / ln(2^32-1)) rounds of a stable sort by fresh 32-bit keys).

Pinned by tests/test_jax_prng.py against the Random123 known-answer vectors (the same three that
jax's own test-suite uses) and against the published outputs of `jax.random.split(PRNGKey(0))`,
`jax.random.uniform(PRNGKey(0))` and `jax.random.normal(PRNGKey(0))`.  `permutation` has no
published vector here: it is pinned through the bit stream it sorts by.

What this buys: `inference.fit_map(..., batch_order='jax')` visits the rows in the order the
reference would for the same `seed`.  It does NOT make whole fits bit-reproducible against the
reference -- the kernel initialisation goes through TFP's TruncatedNormal sampler, which is not
restated.
"""
from __future__ import annotations

import numpy as np
from scipy import special

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
  return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(key, x0, x1):
  """Threefry-2x32-20 of the counter words (x0, x1) under `key` (two uint32 words)."""
  with np.errstate(over='ignore'):
    k0, k1 = _U32(key[0]), _U32(key[1])
    ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
    x0 = np.asarray(x0, dtype=_U32) + ks[0]
    x1 = np.asarray(x1, dtype=_U32) + ks[1]
    for i in range(5):
      for r in _ROT[i % 2]:
        x0 = x0 + x1
        x1 = _rotl(x1, r) ^ x0
      x0 = x0 + ks[(i + 1) % 3]
      x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
  return x0, x1


def _threefry_stream(key, counts):
  """`jax._src.prng.threefry_2x32(key, counts)`: counts is a flat uint32 vector."""
  counts = np.asarray(counts, dtype=_U32).ravel()
  odd = counts.size % 2
  if odd:
    counts = np.concatenate([counts, np.zeros(1, _U32)])
  half = counts.size // 2
  y0, y1 = threefry2x32(key, counts[:half], counts[half:])
  out = np.concatenate([y0, y1])
  return out[:-1] if odd else out


def prng_key(seed) -> np.ndarray:
  """`jax.random.PRNGKey(seed)` (threefry_seed): [high 32 bits, low 32 bits]; a 2-word array
  passes through."""
  a = np.asarray(seed)
  if a.shape == (2,):
    return a.astype(_U32)
  s = int(seed) & 0xFFFFFFFFFFFFFFFF
  return np.array([s >> 32, s & 0xFFFFFFFF], dtype=_U32)


def split(key, num=2) -> np.ndarray:
  """`jax.random.split(key, num)`; `num` may be a shape tuple.  Returns shape + (2,) uint32."""
  shape = (num,) if isinstance(num, (int, np.integer)) else tuple(num)
  n = int(np.prod(shape))
  return _threefry_stream(key, np.arange(2 * n, dtype=_U32)).reshape(shape + (2,))


def fold_in(key, data) -> np.ndarray:
  """`jax.random.fold_in(key, data)` for a 32-bit `data`."""
  return _threefry_stream(key, prng_key(int(data) & 0xFFFFFFFF))


def random_bits(key, n) -> np.ndarray:
  """`jax.random.bits(key, (n,), uint32)`."""
  return _threefry_stream(key, np.arange(n, dtype=_U32))


def uniform(key, n=None, minval=0.0, maxval=1.0) -> np.ndarray:
  """`jax.random.uniform(key, (n,), float32, minval, maxval)`; n=None draws a scalar."""
  bits = random_bits(key, 1 if n is None else n)
  f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
  lo, hi = np.float32(minval), np.float32(maxval)
  out = np.maximum(lo, f * (hi - lo) + lo)
  return out[0] if n is None else out


def normal(key, n=None) -> np.ndarray:
  """`jax.random.normal(key, (n,), float32)`: sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1))."""
  u = uniform(key, n, np.nextafter(np.float32(-1.0), np.float32(0.0)), 1.0)
  return (np.float32(np.sqrt(2.0)) * special.erfinv(u.astype(np.float64))).astype(np.float32)


def permutation(key, n) -> np.ndarray:
  """`jax.random.permutation(key, n)` for an integer n (`_shuffle` of arange(n))."""
  x = np.arange(n)
  key = np.asarray(key, dtype=_U32)
  rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
  for _ in range(rounds):
    key, sub = split(key, 2)
    x = x[np.argsort(random_bits(sub, n), kind='stable')]
  return x


def map_batch_orders(seed, num_devices, members, n_rows, num_epochs, split_index=None) -> np.ndarray:
  """Row order of every (epoch, device, member) of `ensemble_map` (inference.py:571-618):

      seed_i = fold_in(seed, i) if num_splits > 1 else seed          (:436-441)
      _, opt_seed = split(seed_i, 2)                                  (:571)
      member keys = split(opt_seed, (num_devices, members))           (:618)
      per epoch:  key, permute_key = split(key, 2); permutation(permute_key, N)   (:591-595)

  Returns int32 [num_epochs, num_devices, members, n_rows]."""
  key = prng_key(seed)
  if split_index is not None:
    key = fold_in(key, split_index)
  opt_seed = split(key, 2)[1]
  keys = split(opt_seed, (num_devices, members))
  out = np.empty((num_epochs, num_devices, members, n_rows), dtype=np.int32)
  for d in range(num_devices):
    for e in range(members):
      k = keys[d, e]
      for ep in range(num_epochs):
        k, pk = split(k, 2)
        out[ep, d, e] = permutation(pk, n_rows)
  return out
