"""Row data for the hot path: the counterpart of utils/Dataset.scala.

`Data` is the array form of the reference's `Array[(Vec, Int)]` (utils/Dataset.scala:11): CSR with
0-based int32 columns (reference feature key - 1), fp32 values and +/-1 int8 labels.

  * `rcv1(folder, full)` reads the text files the reference reads (utils/Dataset.scala:13-58).
  * `synthetic_rcv1(...)` generates RCV1-shaped rows deterministically from one seed (there is no RCV1
    copy and no network here); the generator is C (csrc/dsgd_host.c) so the full 700 k x 47 236 set takes
    seconds.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .. import native

RCV1_FEATURES = 47236  # utils/Dataset.scala:16


@dataclass
class Data:
    row_ptr: np.ndarray  # int64[n_rows + 1]
    col: np.ndarray      # int32[nnz], 0-based, ascending within a row
    val: np.ndarray      # float32[nnz]
    label: np.ndarray    # int8[n_rows], +1 / -1
    dim: int

    @property
    def n_rows(self) -> int:
        return len(self.row_ptr) - 1

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])

    def __len__(self) -> int:
        return self.n_rows

    def split_at(self, n: int) -> Tuple["Data", "Data"]:
        """`data.splitAt(n)` (Main.scala:52)."""
        n = max(0, min(int(n), self.n_rows))
        cut = int(self.row_ptr[n])
        a = Data(self.row_ptr[:n + 1].copy(), self.col[:cut], self.val[:cut], self.label[:n], self.dim)
        b = Data(self.row_ptr[n:] - cut, self.col[cut:], self.val[cut:], self.label[n:], self.dim)
        return a, b

    def head(self, n: int) -> "Data":
        return self.split_at(n)[0]

    def algorithmic_bytes(self, rows: Optional[np.ndarray] = None) -> int:
        """8*nnz_i + 16 bytes per sample (col id + fp32 value per non-zero; two row
        pointers, the sample index and the label)."""
        lens = np.diff(self.row_ptr)
        if rows is not None:
            lens = lens[np.asarray(rows, dtype=np.int64)]
        return int(8 * lens.sum() + 16 * lens.size)


def synthetic_rcv1(n_rows: int = 700_000, dim: int = RCV1_FEATURES, seed: int = 0, mean_nnz: float = 94.5,
                   sigma: float = 0.7, max_nnz: int = 2000, zipf_s: float = 1.1, zipf_q: float = 20.0,
                   label_noise: float = 0.1, return_w_star: bool = False):
    """RCV1-shaped synthetic rows: density ~0.2 % (mean 94.5 nnz/row, lognormal row
    lengths clipped to [1, 2000]); columns drawn without replacement per row from a Zipf-Mandelbrot
    popularity 1/(rank + q + 1)^s scattered over the id space by a fixed shuffle, sorted ascending;
    values |N(0,1)| row-L2-normalised fp32; labels sign(x.w* + noise) from a planted dense w*."""
    h = native.host_lib()
    p = native.SynthParams(seed, n_rows, dim, mean_nnz, sigma, max_nnz, zipf_s, zipf_q, label_noise)
    row_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    nnz = h.dsgd_synth_row_ptr(C.byref(p), row_ptr.ctypes.data_as(C.c_void_p))
    if nnz < 0:
        raise ValueError("synthetic_rcv1: bad parameters")
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float32)
    label = np.empty(n_rows, dtype=np.int8)
    w_star = np.empty(dim, dtype=np.float64)
    rc = h.dsgd_synth_fill(C.byref(p), row_ptr.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p),
                           val.ctypes.data_as(C.c_void_p), label.ctypes.data_as(C.c_void_p),
                           w_star.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise MemoryError("synthetic_rcv1: generator failed")
    data = Data(row_ptr, col, val, label, dim)
    return (data, w_star) if return_w_star else data


def _read_vectors(path: str, dim: int):
    h = native.host_lib()
    n, nnz = C.c_int64(), C.c_int64()
    if h.dsgd_rcv1_count(path.encode(), C.byref(n), C.byref(nnz)) != 0:
        raise FileNotFoundError(path)
    row_ptr = np.zeros(n.value + 1, dtype=np.int64)
    col = np.empty(nnz.value, dtype=np.int32)
    val = np.empty(nnz.value, dtype=np.float32)
    ids = np.empty(n.value, dtype=np.int64)
    rc = h.dsgd_rcv1_parse(path.encode(), dim, n.value, nnz.value, row_ptr.ctypes.data_as(C.c_void_p),
                           col.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p),
                           ids.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError(f"{path}: malformed RCV1 vectors file (code {rc})")
    return row_ptr, col, val, ids


def rcv1(folder: str, full: bool = True, features_count: int = RCV1_FEATURES) -> Data:
    """utils/Dataset.scala:13-58: train file (+ the four test parts when `full`), labels from the qrels
    file (+1 iff CCAT; the last line of a document wins, quirk Q10)."""
    files = [os.path.join(folder, "lyrl2004_vectors_train.dat")]
    if full:
        files += [os.path.join(folder, f"lyrl2004_vectors_test_pt{d}.dat") for d in range(4)]
    parts = [_read_vectors(f, features_count) for f in files]
    row_ptr = [np.zeros(1, dtype=np.int64)]
    off = 0
    for rp, _, _, _ in parts:
        row_ptr.append(rp[1:] + off)
        off += int(rp[-1])
    row_ptr = np.concatenate(row_ptr)
    col = np.concatenate([p[1] for p in parts])
    val = np.concatenate([p[2] for p in parts])
    ids = np.concatenate([p[3] for p in parts])
    label = np.zeros(len(ids), dtype=np.int8)
    h = native.host_lib()
    qrels = os.path.join(folder, "rcv1-v2.topics.qrels")
    if h.dsgd_rcv1_labels(qrels.encode(), ids.ctypes.data_as(C.c_void_p), len(ids), label.ctypes.data_as(C.c_void_p)) != 0:
        raise FileNotFoundError(qrels)
    if (label == 0).any():
        raise KeyError("rcv1: a document has no qrels line (labels(id) throws in the reference, Dataset.scala:58)")
    return Data(row_ptr, col, val, label, features_count)


def write_rcv1(data: Data, folder: str, first_id: int = 1, name: str = "lyrl2004_vectors_train.dat") -> None:
    """Export rows in the reference's text format so a JVM run of the reference can read the same data."""
    os.makedirs(folder, exist_ok=True)
    h = native.host_lib()
    rc = h.dsgd_rcv1_write(os.path.join(folder, name).encode(), os.path.join(folder, "rcv1-v2.topics.qrels").encode(),
                           data.n_rows, data.row_ptr.ctypes.data_as(C.c_void_p), data.col.ctypes.data_as(C.c_void_p),
                           data.val.ctypes.data_as(C.c_void_p), data.label.ctypes.data_as(C.c_void_p), first_id)
    if rc != 0:
        raise OSError("write_rcv1 failed")
