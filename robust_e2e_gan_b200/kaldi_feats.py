"""Input side of the hot path (next-row N4 of SURVEY.md 8f): Kaldi feature archives -> pinned, length-sorted batches.

Host-side Python/numpy, like the reference's own loader (it is I/O, not arithmetic):

* ``read_mat`` / ``read_mat_scp`` / ``read_mat_ark`` / ``write_mat``: binary Kaldi matrices ('FM ', 'DM ') and the
  one-byte compressed format 'CM ' (data/kaldi_io.py:359-458), ``path:offset`` scp entries (data/kaldi_io.py:36-66).
  Matrices are returned as writable C-contiguous copies (float32; float64 for 'DM ', as in the reference); an .ark
  that is read repeatedly is memory-mapped once (at most 64 mappings are kept, least recently used evicted).
* ``log_spectrum``: the per-utterance transform of MixSequentialDataset.__getitem__
  (data/mix_data_loader.py:198-237): clamp ``<= 1e-7`` -> ``10 log10`` -> input CMVN ``(x + c0) * c1``
  (data/audioparse.py:445-458 with delta_order = 0 and no splicing).
* ``collate``: ``_collate_fn`` (data/mix_data_loader.py:264-302): sort by length (stable, longest first), zero-pad
  into (B, Tmax, F) -- written straight into PINNED host tensors so that ``StepRunner.submit`` / ``.cuda(non_blocking)``
  copies overlap the previous step -- flat int64 targets, int32 sizes.
"""
import mmap
import os
import re
import struct

import numpy as np
import torch

_MAPS = {}          # path -> (mmap, (size, mtime)); insertion ordered, least recently used first
_MAX_MAPS = 64      # open archive mappings kept (address space / file handles stay bounded over a large corpus)


def _open_at(spec):
    """'[ark:]path[:offset]' -> (buffer, position).  Whole files are mapped read-only and cached."""
    if re.search(r'^(ark|scp)(,scp|,b|,t|,n?f|,n?p|,b?o|,n?s|,n?cs)*:', spec):
        spec = spec.split(':', 1)[1]
    offset = 0
    if re.search(r':[0-9]+$', spec):
        spec, off = spec.rsplit(':', 1)
        offset = int(off)
    key = os.path.abspath(spec)
    st = os.stat(key)
    ent = _MAPS.pop(key, None)
    if ent is None or ent[1] != (st.st_size, st.st_mtime_ns):
        with open(key, 'rb') as f:
            buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) if st.st_size else b''
        ent = (buf, (st.st_size, st.st_mtime_ns))
    _MAPS[key] = ent                      # (re)insert as most recently used
    while len(_MAPS) > _MAX_MAPS:         # evict the least recently used mapping; matrices handed out are copies,
        _MAPS.pop(next(iter(_MAPS)))      # so dropping the mmap object is safe (it closes once unreferenced)
    return ent[0], offset


def _decode_compressed(buf, pos):
    """'CM ' one-byte-per-element format (kaldi compressed-matrix.h; data/kaldi_io.py:404-446)."""
    gmin, grange, rows, cols = struct.unpack_from('<ffii', buf, pos)
    pos += 16
    hdr = np.frombuffer(buf, dtype='<u2', count=cols * 4, offset=pos).reshape(cols, 4).astype(np.float32)
    pos += cols * 8
    data = np.frombuffer(buf, dtype=np.uint8, count=rows * cols, offset=pos).reshape(cols, rows)
    pos += rows * cols
    # np.float32(min + range * 1.52590218966964e-05 * value), evaluated left to right in float32 (numpy >= 2 scalar rules)
    p = np.float32(gmin) + (np.float32(grange) * np.float32(1.52590218966964e-05)) * hdr
    p0, p25, p75, p100 = (p[:, i:i + 1] for i in range(4))
    v = data.astype(np.float32)
    lo = p0 + (p25 - p0) / np.float32(64.) * v
    mid = p25 + (p75 - p25) / np.float32(128.) * (v - 64)
    hi = p75 + (p100 - p75) / np.float32(63.) * (v - 192)
    mat = np.where(data <= 64, lo, np.where(data <= 192, mid, hi)).astype(np.float32)
    return np.ascontiguousarray(mat.T), pos


def _parse_mat(buf, pos):
    """Binary matrix starting at the '\\0B' marker.  Returns (C-contiguous WRITABLE array -- a copy, never a view of
    the read-only mapping; float32 for 'FM ' / 'CM ', float64 for 'DM ' exactly as the reference's reader returns
    them, data/kaldi_io.py:388-397 -- and the position after the matrix)."""
    if bytes(buf[pos:pos + 2]) != b'\0B':
        raise ValueError("kaldi_feats: only binary Kaldi matrices are supported (missing \\0B marker)")
    header = bytes(buf[pos + 2:pos + 5]).decode()
    pos += 5
    if header.startswith('CM'):
        if header != 'CM ':
            raise ValueError("kaldi_feats: compressed formats CM2/CM3 are not supported (as in the reference)")
        return _decode_compressed(buf, pos)
    if header == 'FM ':
        dt = np.dtype('<f4')
    elif header == 'DM ':
        dt = np.dtype('<f8')
    else:
        raise ValueError("kaldi_feats: unknown matrix header %r" % header)
    s1, rows, s2, cols = struct.unpack_from('<bibi', buf, pos)
    pos += 10
    if s1 != 4 or s2 != 4 or rows < 0 or cols < 0:
        raise ValueError("kaldi_feats: corrupt matrix dimensions")
    n = rows * cols
    mat = np.array(np.frombuffer(buf, dtype=dt, count=n, offset=pos).reshape(rows, cols), order='C')
    return mat, pos + n * dt.itemsize


def read_mat(spec):
    """One matrix from 'path' or 'path:offset' (an scp value)."""
    buf, pos = _open_at(spec)
    return _parse_mat(buf, pos)[0]


def read_mat_ark(path):
    """Iterate (key, matrix) over a binary .ark file."""
    buf, pos = _open_at(path)
    n = len(buf)
    while pos < n:
        end = buf.find(b' ', pos)
        if end < 0:
            break
        key = bytes(buf[pos:end]).decode('latin1').strip()
        if not key:
            break
        mat, pos = _parse_mat(buf, end + 1)
        yield key, mat


def read_scp(path):
    """[(key, 'ark_path:offset'), ...] of an .scp file."""
    out = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                key, val = line.split(None, 1)
                out.append((key, val))
    return out


def read_mat_scp(path):
    for key, val in read_scp(path):
        yield key, read_mat(val)


def write_mat(fd, m, key=''):
    """Binary 'FM ' / 'DM ' matrix (data/kaldi_io.py:461-491); returns the offset of the matrix for an scp entry."""
    m = np.ascontiguousarray(m)
    if key:
        fd.write((key + ' ').encode('latin1'))
    off = fd.tell()
    fd.write(b'\0B')
    if m.dtype == np.float32:
        fd.write(b'FM ')
    elif m.dtype == np.float64:
        fd.write(b'DM ')
    else:
        raise ValueError("kaldi_feats.write_mat: float32 or float64 only, got %s" % m.dtype)
    fd.write(b'\x04' + struct.pack('<I', m.shape[0]) + b'\x04' + struct.pack('<I', m.shape[1]))
    fd.write(m.tobytes())
    return off


def log_spectrum(spect, cmvn=None):
    """data/mix_data_loader.py:203-207: ``spect[spect <= 1e-7] = 1e-7`` (IN PLACE, as the reference does: the clamped
    magnitudes are what the batch carries), ``10 log10``, then input CMVN (audioparse.py:452-453)."""
    spect[spect <= 1e-7] = 1e-7
    out = 10 * np.log10(spect)
    if cmvn is not None:
        out = (out + cmvn[0, :]) * cmvn[1, :]
    return out


def _pinned(shape, dtype, pin):
    t = torch.zeros(shape, dtype=dtype)
    return t.pin_memory() if pin else t


def collate(batch, pin=True):
    """``_collate_fn`` (data/mix_data_loader.py:264-302) into pinned tensors.  ``batch`` is a list of samples
    (utt_id, spk_id, clean_spect, clean_log_spect, mix_spect, mix_log_spect, cos_angle, target) whose arrays may be
    numpy or torch.  Returns the reference's 10-tuple."""
    pin = pin and torch.cuda.is_available()
    batch = sorted(batch, key=lambda s: s[2].shape[0], reverse=True)
    B = len(batch)
    Tmax, F = batch[0][2].shape[0], batch[0][2].shape[1]
    outs = [_pinned((B, Tmax, F), torch.float32, pin) for _ in range(5)]
    input_sizes = _pinned((B,), torch.int32, pin)
    target_sizes = _pinned((B,), torch.int32, pin)
    targets, utt_ids, spk_ids = [], [], []
    for x, s in enumerate(batch):
        utt_ids.append(s[0])
        spk_ids.append(s[1])
        L = s[2].shape[0]
        for dst, src in zip(outs, s[2:7]):
            dst[x, :L].copy_(torch.as_tensor(np.asarray(src), dtype=torch.float32) if not torch.is_tensor(src) else src)
        input_sizes[x] = L
        target_sizes[x] = len(s[7])
        targets.extend(int(t) for t in s[7])
    targets = torch.tensor(targets, dtype=torch.int64)
    clean, clean_log, mix, mix_log, cos = outs
    return utt_ids, spk_ids, clean, clean_log, mix, mix_log, cos, targets, input_sizes, target_sizes
