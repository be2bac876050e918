"""Device-resident corpus + on-GPU batch construction: the drop-in for `DatasetBuilder.refresh_train_dataset` /
`build_data` + the `DataLoader(shuffle=True)` of the reference's epoch loop (model/dataset_builder.py:55-63, :112-204,
main.py:160-169) for the method-name task (`build`, `epoch`) and the variable-name task (`build_vars`, `epoch_vars`).  All arithmetic happens in libc2v_b200.so (`c2v_build_batch`); there is no
CPU fallback.

    corpus = DeviceCorpus.from_reader(reader, builder.train_items, device)          # once
    for starts, paths, ends, label in corpus.epoch(batch_size, max_path_length, seed=epoch):
        preds, _, _ = model.forward(starts, paths, ends, label)                      # main.py:172
"""
import ctypes

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class DeviceCorpus:
    """CSR image of `reader.items` in HBM: offsets int64 [n+1], contexts int32 [total, 3], labels int64 [n]."""

    def __init__(self, offsets, contexts, labels, method_token, question_token, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.C2VError("DeviceCorpus lives on a CUDA device (sm_100a); there is no CPU path")
        self.offsets = torch.as_tensor(offsets, dtype=torch.int64).contiguous().to(dev)
        self.contexts = torch.as_tensor(contexts, dtype=torch.int32).reshape(-1, 3).contiguous().to(dev)
        self.labels = None if labels is None else torch.as_tensor(labels, dtype=torch.int64).contiguous().to(dev)
        self.n_items = int(self.offsets.numel() - 1)
        if self.n_items < 1 or int(self.offsets[-1]) != self.contexts.shape[0]:
            raise ValueError("offsets / contexts do not describe a CSR corpus")
        self.method_token, self.question_token = int(method_token), int(question_token)
        self.device = dev

    @classmethod
    def from_reader(cls, reader, items, device):
        """`reader`: the reference's DatasetReader (dataset_reader.py:44-128); `items`: e.g. builder.train_items."""
        import numpy as np
        off = np.zeros(len(items) + 1, dtype=np.int64)
        for i, it in enumerate(items):
            off[i + 1] = off[i] + len(it.path_contexts)
        ctx = np.asarray([pc for it in items for pc in it.path_contexts], dtype=np.int32).reshape(-1, 3)
        lab = np.asarray([reader.label_vocab.stoi[it.normalized_label] for it in items], dtype=np.int64)
        return cls(off, ctx, lab, reader.terminal_vocab.stoi["@method_0"], reader.QUESTION_TOKEN_INDEX, device)

    @classmethod
    def from_corpus(cls, reader, device, item_indices=None):
        """`reader`: code2vec_b200.corpus.CorpusReader (the C++ parser's CSR arrays go to HBM as they are);
        item_indices: subset / order of items (e.g. the train split), default all."""
        import numpy as np
        if item_indices is None:
            off, ctx, lab = reader.ctx_offsets, reader.contexts, reader.item_labels
        else:
            idx = np.asarray(item_indices, dtype=np.int64)
            n = (reader.ctx_offsets[idx + 1] - reader.ctx_offsets[idx])
            off = np.zeros(len(idx) + 1, np.int64); np.cumsum(n, out=off[1:])
            take = np.concatenate([np.arange(reader.ctx_offsets[i], reader.ctx_offsets[i + 1]) for i in idx]) if len(idx) else np.zeros(0, np.int64)
            ctx, lab = reader.contexts[take], reader.item_labels[idx]
        c = cls(off, ctx, lab, reader.terminal_vocab.stoi["@method_0"], reader.QUESTION_TOKEN_INDEX, device)
        if reader.infer_variable:
            ui, uv, ul = reader.variable_units(item_indices)
            c.set_variable_units(ui, uv, ul, reader.variable_indexes, reader.terminal_vocab.len(),
                                 reader.shuffle_variable_indexes)
        return c

    def set_variable_units(self, unit_item, unit_var, unit_label, variable_indexes, terminal_count, shuffle_variable_indexes=False):
        """The bags of the variable-name task (dataset_builder.py:152-204): unit u = (item unit_item[u] of THIS corpus,
        terminal index unit_var[u] of its @var alias, label unit_label[u])."""
        import numpy as np
        dev = self.device
        self.unit_item = torch.as_tensor(np.asarray(unit_item), dtype=torch.int64).contiguous().to(dev)
        self.unit_var = torch.as_tensor(np.asarray(unit_var), dtype=torch.int64).contiguous().to(dev)
        self.unit_label = torch.as_tensor(np.asarray(unit_label), dtype=torch.int64).contiguous().to(dev)
        self.n_units = int(self.unit_item.numel())
        var = np.asarray(variable_indexes, dtype=np.int64)
        pos = np.full(int(terminal_count), -1, np.int32)
        pos[var] = np.arange(len(var), dtype=np.int32)
        self.var_pos = torch.from_numpy(pos).to(dev)
        self.variable_indexes = torch.from_numpy(var).to(dev)
        self.terminal_count, self.shuffle_variable_indexes = int(terminal_count), bool(shuffle_variable_indexes)

    def build_vars(self, unit_ids, max_path_length, seed):
        """-> (starts, paths, ends, label) of the variable-name bags `unit_ids` (int64 [B] into the units)."""
        lib = _lib.load()
        if getattr(self, "unit_item", None) is None:
            raise ValueError("no variable units: build the corpus from a reader with infer_variable=True")
        ids = unit_ids.to(device=self.device, dtype=torch.int64).contiguous()
        B, L = int(ids.numel()), int(max_path_length)
        with torch.cuda.device(self.device):
            starts = torch.empty((B, L), dtype=torch.int64, device=self.device)
            paths = torch.empty_like(starts); ends = torch.empty_like(starts)
            label = torch.empty((B,), dtype=torch.int64, device=self.device)
            rc = lib.c2v_build_batch_vars(_ptr(self.offsets), _ptr(self.contexts), self.n_items, _ptr(self.unit_item),
                                          _ptr(self.unit_var), _ptr(self.unit_label), self.n_units, _ptr(ids), B, L,
                                          int(seed) & 0xFFFFFFFFFFFFFFFF, self.question_token, _ptr(self.var_pos),
                                          self.terminal_count, _ptr(self.variable_indexes), int(self.variable_indexes.numel()),
                                          1 if self.shuffle_variable_indexes else 0,
                                          _ptr(starts), _ptr(paths), _ptr(ends), _ptr(label),
                                          ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            _lib.check(rc, "c2v_build_batch_vars")
        return starts, paths, ends, label

    def build(self, item_ids, max_path_length, seed, check=False):
        """-> (starts, paths, ends, label): int64 [B, L] x3 and [B], like `build_data` + the DataLoader collate."""
        lib = _lib.load()
        ids = item_ids.to(device=self.device, dtype=torch.int64).contiguous()
        if check and ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= self.n_items):
            raise IndexError("item id out of range")
        B, L = int(ids.numel()), int(max_path_length)
        with torch.cuda.device(self.device):
            starts = torch.empty((B, L), dtype=torch.int64, device=self.device)
            paths = torch.empty_like(starts); ends = torch.empty_like(starts)
            label = torch.empty((B,), dtype=torch.int64, device=self.device)
            rc = lib.c2v_build_batch(_ptr(self.offsets), _ptr(self.contexts), self.n_items, _ptr(ids), _ptr(self.labels),
                                     B, L, int(seed) & 0xFFFFFFFFFFFFFFFF, self.method_token, self.question_token,
                                     _ptr(starts), _ptr(paths), _ptr(ends), _ptr(label),
                                     ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            _lib.check(rc, "c2v_build_batch")
        return starts, paths, ends, label

    def epoch(self, batch_size, max_path_length, seed, shuffle=True, rank=0, world=1):
        """One pass over the corpus in random order (main.py:160-162); with world > 1 each rank takes its strided shard
        of the same permutation.  Every (epoch seed, item) pair draws a fresh context subset, like the per-epoch
        `refresh_train_dataset` of the reference."""
        g = torch.Generator(device=self.device).manual_seed(int(seed))
        order = torch.randperm(self.n_items, generator=g, device=self.device) if shuffle else \
            torch.arange(self.n_items, device=self.device)
        order = order[rank::world]
        for lo in range(0, order.numel(), batch_size):          # last batch ragged (drop_last unset, main.py:162)
            yield self.build(order[lo:lo + batch_size], max_path_length, seed)

    def epoch_vars(self, batch_size, max_path_length, seed, shuffle=True, rank=0, world=1):
        """the same pass over the variable-name units (dataset_builder.py:152-204)"""
        g = torch.Generator(device=self.device).manual_seed(int(seed))
        order = torch.randperm(self.n_units, generator=g, device=self.device) if shuffle else \
            torch.arange(self.n_units, device=self.device)
        order = order[rank::world]
        for lo in range(0, order.numel(), batch_size):
            yield self.build_vars(order[lo:lo + batch_size], max_path_length, seed)
