"""CPU restatement (numpy) of the reference's per-epoch batch construction for the method-name task,
`DatasetBuilder.build_data` (/root/reference/model/dataset_builder.py:112-147, infer_method branch) --
TEST INFRASTRUCTURE ONLY: the checker of `c2v_build_batch` (code2vec_b200/csrc/c2v_batch.cu), never the product path.

Reference semantics, per item: `random.shuffle(item.path_contexts)`; take the first `max_path_length`; rewrite a
start / end equal to `@method_0` to `@question` (:136-143); pad the three lists with 0 up to `max_path_length`
(:145-147, :212-219).  "Shuffle, then take the first L" is a uniformly random subset of size min(n, L) in uniformly
random order.  Python's Mersenne-Twister stream cannot be reproduced on the GPU, so -- exactly like dropout -- the
random choice is re-derived from a counter-based hash and parity is defined on what does not depend on the stream:

  * key(j) = high 32 bits of splitmix64(seed, item, j) for context j of the item;
  * n <= L: all n contexts, in stored order (the model is invariant to the order inside a bag: softmax and the
    weighted sum are symmetric; tests check that);
  * n >  L: the L contexts with the smallest (key, j), in stored order -- iid keys make this a uniform subset.

The kernel must reproduce this function bit for bit; this function is pinned to the reference by
tests/golden/builder_corpus.npz (the reference's own reader + builder on dataset/corpus.txt): same bag sizes, every
bag a sub-multiset of the item's rewritten contexts, identical multisets whenever n <= L, identical labels.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    """splitmix64 finalizer (Steele, Lea, Flood 2014) on uint64 arrays."""
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        x = x ^ (x >> np.uint64(31))
    return x


def context_keys(seed, item, n):
    """uint32 selection key of every context j < n of `item` under `seed`."""
    with np.errstate(over="ignore"):
        base = _mix64(np.array([np.uint64(seed) ^ (np.uint64(item) * np.uint64(0xD1B54A32D192ED03) & _M64)], dtype=np.uint64))[0]
        j = np.arange(n, dtype=np.uint64)
        k = _mix64(base ^ (j * np.uint64(0x8CB92BA72F3D8DD7) & _M64))
    return (k >> np.uint64(32)).astype(np.uint32)


def select(seed, item, n, L):
    """indices (ascending) of the contexts of `item` that enter the bag."""
    if n <= L:
        return np.arange(n, dtype=np.int64)
    keys = context_keys(seed, item, n)
    order = np.lexsort((np.arange(n), keys))           # by key, ties by index
    return np.sort(order[:L]).astype(np.int64)


def build_batch(offsets, contexts, item_ids, L, seed, method_token, question_token):
    """-> starts, paths, ends int64 [B, L] (dataset_builder.py:127-150, :206-208)."""
    B = len(item_ids)
    out = np.zeros((3, B, L), dtype=np.int64)
    for b, item in enumerate(np.asarray(item_ids, dtype=np.int64)):
        lo, hi = int(offsets[item]), int(offsets[item + 1])
        sel = select(seed, int(item), hi - lo, L)
        c = np.asarray(contexts[lo:hi], dtype=np.int64)[sel]
        s, p, e = c[:, 0].copy(), c[:, 1], c[:, 2].copy()
        s[s == method_token] = question_token          # :136-137
        e[e == method_token] = question_token          # :142-143
        out[0, b, :len(sel)] = s; out[1, b, :len(sel)] = p; out[2, b, :len(sel)] = e
    return out[0], out[1], out[2]


# ---- variable-name task (dataset_builder.py:152-204, infer_variable branch) -------------------------------------------
def var_permutation(seed, item, n_vars):
    """sigma: position i of `variable_indexes` is replaced by variable_indexes[sigma[i]] (:166-168 shuffle, re-derived
    from a counter-based hash like everything else here): argsort of key(i) = hash(seed ^ A5.., item, i), ties by i."""
    with np.errstate(over="ignore"):
        vbase = _mix64(np.array([np.uint64(seed) ^ np.uint64(0xA5A5A5A55A5A5A5A) ^
                                 (np.uint64(item) * np.uint64(0xD1B54A32D192ED03) & _M64)], dtype=np.uint64))[0]
        i = np.arange(n_vars, dtype=np.uint64)
        k = (_mix64(vbase ^ (i * np.uint64(0x8CB92BA72F3D8DD7) & _M64)) >> np.uint64(32)).astype(np.uint32)
    return np.lexsort((np.arange(n_vars), k)).astype(np.int64)


def build_batch_vars(offsets, contexts, unit_item, unit_var, unit_ids, L, seed, question_token, variable_indexes=None,
                     shuffle_variable_indexes=False):
    """-> starts, paths, ends int64 [B, L]: row b = the contexts of item unit_item[u] that touch token unit_var[u]
    (u = unit_ids[b]), @var -> @question, other variables permuted per item if asked, min(n, L) of them chosen by the
    smallest (key, j) with key = hash(seed, item, var, j), kept in stored order, zero suffix."""
    B = len(unit_ids)
    out = np.zeros((3, B, L), dtype=np.int64)
    var = None if variable_indexes is None else np.asarray(variable_indexes, dtype=np.int64)
    for b, u in enumerate(np.asarray(unit_ids, dtype=np.int64)):
        item, v = int(unit_item[u]), int(unit_var[u])
        lo, hi = int(offsets[item]), int(offsets[item + 1])
        c = np.asarray(contexts[lo:hi], dtype=np.int64)
        m = np.nonzero((c[:, 0] == v) | (c[:, 2] == v))[0]
        if len(m) > L:
            with np.errstate(over="ignore"):
                base = _mix64(np.array([np.uint64(seed) ^ (np.uint64(item) * np.uint64(0xD1B54A32D192ED03) & _M64) ^
                                        (np.uint64(v) * np.uint64(0x9E3779B97F4A7C15) & _M64)], dtype=np.uint64))[0]
                k = (_mix64(base ^ (m.astype(np.uint64) * np.uint64(0x8CB92BA72F3D8DD7) & _M64)) >> np.uint64(32)).astype(np.uint32)
            m = np.sort(m[np.lexsort((m, k))[:L]])
        c = c[m]
        s, p, e = c[:, 0].copy(), c[:, 1], c[:, 2].copy()
        if shuffle_variable_indexes and var is not None and len(var) > 1:
            sigma = var_permutation(seed, item, len(var))
            mp = {int(var[i]): int(var[sigma[i]]) for i in range(len(var))}
            s = np.asarray([question_token if t == v else mp.get(int(t), int(t)) for t in s], dtype=np.int64)
            e = np.asarray([question_token if t == v else mp.get(int(t), int(t)) for t in e], dtype=np.int64)
        else:
            s[s == v] = question_token
            e[e == v] = question_token
        out[0, b, :len(m)] = s; out[1, b, :len(m)] = p; out[2, b, :len(m)] = e
    return out[0], out[1], out[2]
