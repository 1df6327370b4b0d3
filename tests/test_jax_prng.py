"""bayesnf_b200.jax_prng: known-answer tests.

The three Threefry-2x32-20 vectors are Random123's kat_vectors (also used by jax's own
tests/random_test.py::testThreefry2x32); the split / uniform / normal values are the outputs
jax documents for PRNGKey(0) with the default threefry2x32 implementation."""
import numpy as np

from bayesnf_b200 import jax_prng as J


def test_threefry2x32_known_answers():
  def run(key, ctr):
    y0, y1 = J.threefry2x32(np.array(key, np.uint32), np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
    return int(y0[0]), int(y1[0])
  assert run((0, 0), (0, 0)) == (0x6B200159, 0x99BA4EFE)
  assert run((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF)) == (0x1CB996FC, 0xBB002BE7)
  assert run((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3)) == (0xC4923A9C, 0x483DF7A0)


def test_key_and_split():
  np.testing.assert_array_equal(J.prng_key(0), [0, 0])
  np.testing.assert_array_equal(J.prng_key(42), [0, 42])
  np.testing.assert_array_equal(J.prng_key((1 << 32) + 5), [1, 5])
  np.testing.assert_array_equal(J.split(J.prng_key(0)), [[4146024105, 967050713], [2718843009, 1272950319]])
  s = J.split(J.prng_key(7), (2, 3))
  assert s.shape == (2, 3, 2) and len({tuple(k) for k in s.reshape(-1, 2)}) == 6
  np.testing.assert_array_equal(s.reshape(-1, 2), J.split(J.prng_key(7), 6))      # same counter layout
  assert not np.array_equal(J.fold_in(J.prng_key(0), 1), J.fold_in(J.prng_key(0), 2))


def test_uniform_and_normal_of_key0():
  assert abs(float(J.uniform(J.prng_key(0))) - 0.41845703) < 1e-7
  assert abs(float(J.normal(J.prng_key(0))) - (-0.20584226)) < 1e-6
  u = J.uniform(J.prng_key(3), 10001)
  assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 0.02
  z = J.normal(J.prng_key(3), 20000)
  assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03


def test_permutation_is_a_stable_sort_by_the_bit_stream():
  key = J.prng_key(11)
  p = J.permutation(key, 100)
  assert sorted(p.tolist()) == list(range(100))
  _, sub = J.split(key, 2)                                 # one round for n = 100
  np.testing.assert_array_equal(p, np.argsort(J.random_bits(sub, 100), kind='stable'))
  big = J.permutation(key, 5000)                           # ceil(3 ln 5000 / ln(2^32-1)) = 2 rounds
  assert sorted(big.tolist()) == list(range(5000)) and not np.array_equal(big[:100], p)


def test_map_batch_orders_follow_the_reference_key_tree():
  o = J.map_batch_orders(np.array([0, 5], np.uint32), 1, 3, 50, 4)
  assert o.shape == (4, 1, 3, 50) and o.dtype == np.int32
  for ep in range(4):
    for e in range(3):
      assert sorted(o[ep, 0, e].tolist()) == list(range(50))
  assert not np.array_equal(o[0, 0, 0], o[0, 0, 1]) and not np.array_equal(o[0, 0, 0], o[1, 0, 0])
  # member 1, epoch 1, by hand
  opt = J.split(J.prng_key(5), 2)[1]
  k = J.split(opt, (1, 3))[0, 1]
  k, _ = J.split(k, 2)
  _, pk = J.split(k, 2)
  np.testing.assert_array_equal(o[1, 0, 1], J.permutation(pk, 50))
  with_split = J.map_batch_orders(5, 1, 3, 50, 1, split_index=2)
  assert not np.array_equal(with_split[0], o[0])
