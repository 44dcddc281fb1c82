"""Test-only FLAC decoder (pure Python + numpy) for the reference's fixtures.

The container has no libsndfile / soundfile / ffmpeg, and the reference ships its
only real outputs as 24-bit FLAC (`/root/reference/subtraction_demo/*.flac`, written
by `util_audio.audio_to_flac`, `/root/reference/util_audio.py:966-968`).  This module
decodes exactly the subset libFLAC emits (CONSTANT / VERBATIM / FIXED / LPC subframes,
Rice-coded residuals with escape partitions, all four stereo modes), checks every
frame's CRC-16 and the STREAMINFO MD5 of the decoded PCM, and is used by the oracle
pin tests only -- the product never imports it.

Format notes follow the public FLAC specification (RFC 9639).
"""
import hashlib

import numpy as np


class FlacError(ValueError):
    pass


def _crc_table(poly, bits):
    top = 1 << (bits - 1)
    mask = (1 << bits) - 1
    tab = []
    for b in range(256):
        c = b << (bits - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
        tab.append(c)
    return tab


_CRC8 = _crc_table(0x07, 8)
_CRC16 = _crc_table(0x8005, 16)


def _crc8(data):
    c = 0
    for b in data:
        c = _CRC8[c ^ b]
    return c


def _crc16(data):
    c = 0
    for b in data:
        c = ((c << 8) & 0xFFFF) ^ _CRC16[(c >> 8) ^ b]
    return c


class _Bits:
    """MSB-first bit reader over a bytes object, with an index of the next set bit
    (unary Rice prefixes are then one table look-up)."""

    def __init__(self, data):
        self.data = data
        self.pos = 0
        bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8))
        idx = np.where(bits == 1, np.arange(bits.size, dtype=np.int64), np.int64(1) << 60)
        self.next_one = np.minimum.accumulate(idx[::-1])[::-1]

    def read(self, n):
        if n == 0:
            return 0
        p = self.pos
        b0 = p >> 3
        nbytes = ((p & 7) + n + 7) >> 3
        chunk = int.from_bytes(self.data[b0:b0 + nbytes], "big")
        if b0 + nbytes > len(self.data):
            raise FlacError("read past the end of the stream")
        self.pos = p + n
        return (chunk >> (nbytes * 8 - (p & 7) - n)) & ((1 << n) - 1)

    def read_signed(self, n):
        v = self.read(n)
        return v - (1 << n) if v >> (n - 1) else v

    def unary(self):
        p = self.pos
        q = int(self.next_one[p]) - p
        if q < 0 or q > (1 << 40):
            raise FlacError("unterminated unary code")
        self.pos = p + q + 1
        return q

    def align(self):
        self.pos = (self.pos + 7) & ~7


def _read_residual(br, blocksize, order, out):
    method = br.read(2)
    if method > 1:
        raise FlacError("reserved residual coding method")
    pbits = 4 if method == 0 else 5
    esc = (1 << pbits) - 1
    porder = br.read(4)
    nparts = 1 << porder
    if blocksize % nparts:
        raise FlacError("partition order does not divide the block")
    i = order
    for part in range(nparts):
        count = (blocksize >> porder) - (order if part == 0 else 0)
        k = br.read(pbits)
        if k == esc:
            raw = br.read(5)
            for _ in range(count):
                out[i] = br.read_signed(raw) if raw else 0
                i += 1
        else:
            unary, read = br.unary, br.read
            for _ in range(count):
                u = (unary() << k) | read(k)
                out[i] = (u >> 1) ^ -(u & 1)
                i += 1


def _read_subframe(br, blocksize, bps):
    if br.read(1):
        raise FlacError("subframe padding bit set")
    kind = br.read(6)
    wasted = 0
    if br.read(1):
        wasted = br.unary() + 1
        bps -= wasted
    s = [0] * blocksize
    if kind == 0:                                   # CONSTANT
        s = [br.read_signed(bps)] * blocksize
    elif kind == 1:                                 # VERBATIM
        s = [br.read_signed(bps) for _ in range(blocksize)]
    elif 8 <= kind <= 12:                           # FIXED, order kind-8
        order = kind - 8
        for i in range(order):
            s[i] = br.read_signed(bps)
        _read_residual(br, blocksize, order, s)
        if order == 1:
            for i in range(1, blocksize):
                s[i] += s[i - 1]
        elif order == 2:
            for i in range(2, blocksize):
                s[i] += 2 * s[i - 1] - s[i - 2]
        elif order == 3:
            for i in range(3, blocksize):
                s[i] += 3 * s[i - 1] - 3 * s[i - 2] + s[i - 3]
        elif order == 4:
            for i in range(4, blocksize):
                s[i] += 4 * s[i - 1] - 6 * s[i - 2] + 4 * s[i - 3] - s[i - 4]
    elif kind >= 32:                                # LPC, order (kind & 31) + 1
        order = (kind & 31) + 1
        for i in range(order):
            s[i] = br.read_signed(bps)
        prec = br.read(4) + 1
        if prec == 16:
            raise FlacError("reserved LPC precision")
        shift = br.read_signed(5)
        if shift < 0:
            raise FlacError("negative LPC shift")
        coef = [br.read_signed(prec) for _ in range(order)]   # coef[j] multiplies s[i-1-j]
        _read_residual(br, blocksize, order, s)
        rc = coef[::-1]
        for i in range(order, blocksize):
            acc = 0
            w = s[i - order:i]
            for c, x in zip(rc, w):
                acc += c * x
            s[i] += acc >> shift                    # arithmetic shift = floor, as libFLAC
    else:
        raise FlacError("reserved subframe type %d" % kind)
    if wasted:
        s = [x << wasted for x in s]
    return s


_BLOCKSIZES = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608}
_SAMPLESIZES = {1: 8, 2: 12, 4: 16, 5: 20, 6: 24, 7: 32}


def read_flac(path, verify=True):
    """Decode a FLAC file.  Returns (pcm int32 [samples] or [samples, channels],
    sample_rate, bits_per_sample).  With verify=True every frame CRC and the
    STREAMINFO MD5 signature are checked (a mismatch raises FlacError)."""
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"fLaC":
        raise FlacError("not a FLAC stream")
    pos = 4
    info = None
    while True:
        hdr = data[pos]
        length = int.from_bytes(data[pos + 1:pos + 4], "big")
        body = data[pos + 4:pos + 4 + length]
        if (hdr & 0x7F) == 0:
            v = int.from_bytes(body[10:18], "big")
            info = {"sr": v >> 44, "channels": ((v >> 41) & 7) + 1, "bps": ((v >> 36) & 31) + 1,
                    "total": v & ((1 << 36) - 1), "md5": body[18:34]}
        pos += 4 + length
        if hdr & 0x80:
            break
    if info is None:
        raise FlacError("no STREAMINFO block")
    nch, bps_stream = info["channels"], info["bps"]
    br = _Bits(data)
    br.pos = pos * 8
    chans = [[] for _ in range(nch)]
    done = 0
    while (br.pos >> 3) < len(data) and (info["total"] == 0 or done < info["total"]):
        start = br.pos >> 3
        if br.read(14) != 0x3FFE:
            raise FlacError("lost frame sync at byte %d" % start)
        br.read(1)
        br.read(1)                                   # blocking strategy (only affects the coded number)
        bs_code, sr_code = br.read(4), br.read(4)
        ch_code, ss_code = br.read(4), br.read(3)
        br.read(1)
        first = br.read(8)                           # UTF-8 style frame / sample number
        extra = 0
        while first & (0x80 >> extra):
            extra += 1
        for _ in range(max(extra - 1, 0)):
            br.read(8)
        if bs_code == 6:
            blocksize = br.read(8) + 1
        elif bs_code == 7:
            blocksize = br.read(16) + 1
        elif bs_code >= 8:
            blocksize = 256 << (bs_code - 8)
        elif bs_code in _BLOCKSIZES:
            blocksize = _BLOCKSIZES[bs_code]
        else:
            raise FlacError("reserved block size code")
        if sr_code == 12:
            br.read(8)
        elif sr_code in (13, 14):
            br.read(16)
        hdr_end = br.pos >> 3
        crc8 = br.read(8)
        if verify and _crc8(data[start:hdr_end]) != crc8:
            raise FlacError("frame header CRC mismatch")
        bps = _SAMPLESIZES.get(ss_code, bps_stream)
        if ch_code < 8:
            if ch_code + 1 != nch:
                raise FlacError("channel count changed mid-stream")
            subs = [_read_subframe(br, blocksize, bps) for _ in range(nch)]
        elif ch_code == 8:                           # left / side
            l = _read_subframe(br, blocksize, bps)
            sd = _read_subframe(br, blocksize, bps + 1)
            subs = [l, [a - b for a, b in zip(l, sd)]]
        elif ch_code == 9:                           # side / right
            sd = _read_subframe(br, blocksize, bps + 1)
            r = _read_subframe(br, blocksize, bps)
            subs = [[a + b for a, b in zip(sd, r)], r]
        elif ch_code == 10:                          # mid / side
            m = _read_subframe(br, blocksize, bps)
            sd = _read_subframe(br, blocksize, bps + 1)
            l, r = [], []
            for a, b in zip(m, sd):
                a = (a << 1) | (b & 1)
                l.append((a + b) >> 1)
                r.append((a - b) >> 1)
            subs = [l, r]
        else:
            raise FlacError("reserved channel assignment")
        br.align()
        body_end = br.pos >> 3
        crc16 = br.read(16)
        if verify and _crc16(data[start:body_end]) != crc16:
            raise FlacError("frame CRC-16 mismatch")
        for c in range(nch):
            chans[c].extend(subs[c])
        done += blocksize
    pcm = np.asarray(chans, dtype=np.int64).T
    if info["total"]:
        pcm = pcm[:info["total"]]
    if verify and any(info["md5"]):
        nbytes = (bps_stream + 7) // 8
        raw = (pcm.astype("<i8").reshape(-1, 1).view(np.uint8).reshape(-1, 8)[:, :nbytes]).tobytes()
        if hashlib.md5(raw).digest() != info["md5"]:
            raise FlacError("decoded PCM does not match the STREAMINFO MD5")
    pcm = pcm.astype(np.int32)
    return (pcm[:, 0] if nch == 1 else pcm), info["sr"], bps_stream


def pcm_to_float(pcm, bps):
    """libsndfile's integer -> float normalisation (what librosa.load / soundfile.read
    hand the reference): x / 2^(bps-1)."""
    return (pcm.astype(np.float64) / float(1 << (bps - 1))).astype(np.float32)
