"""The confounder draw of src/models/DCCF.py:72 — `torch.randint(item_num, size=(P, S))` on the torch CPU generator —
served by dccf_confounder_draw (csrc/confounder_draw.cu): the same numbers and the same generator state afterwards,
2-3x faster than torch's element-at-a-time loop, which out-lasts the GPU scorer during an evaluation pass (10 draws per
scored row).  The generator's words travel through torch.get_rng_state() / set_rng_state(), so every other consumer of
the torch generator (initialisation, later draws, the reference's own calls) sees exactly the stream it would have seen.

Layout of the state blob (torch CPUGeneratorImpl::get_state, struct CPUGeneratorImplStateLegacy): the_initial_seed u64 @0,
left i32 @8, seeded i32 @12, next u64 @16, state u64[624] @24 (one 32-bit word per slot), Gaussian caches behind.  The
layout is checked once per process against torch.randint itself; if this torch build serialises differently, or the
range is outside torch's 32-bit path (item_num >= 2^28), the draw is torch.randint — the reference's own call.
"""
import ctypes
import threading

import numpy as np
import torch

_STATE_BYTES = 5056
_OFF_LEFT, _OFF_NEXT, _OFF_WORDS, _N_WORDS = 8, 16, 24, 624
_MAX_HIGH = 1 << 28        # torch 2.11 draws 64-bit words for ranges >= 2^28 (uniform_int_from_to_distribution)
MIN_DRAWS = 16384           # below this the two state copies cost more than torch's loop
_checked = None             # None = not yet, True = layout verified, False = unusable
_lock = threading.Lock()


def _native_into(out, high, generator):
    from . import _lib
    lib = _lib.load()
    blob = generator.get_state() if generator is not None else torch.get_rng_state()
    if blob.numel() != _STATE_BYTES:
        return False
    raw = blob.numpy()
    words = raw[_OFF_WORDS:_OFF_WORDS + 8 * _N_WORDS].view(np.uint64)
    state = np.ascontiguousarray(words, dtype=np.uint32)
    left = ctypes.c_int32(int(raw[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0]))
    nxt = ctypes.c_int32(0)
    if not 1 <= left.value <= _N_WORDS:
        return False
    flat = out.view(-1)
    rc = lib.dccf_confounder_draw(state.ctypes.data_as(ctypes.c_void_p), ctypes.byref(left), ctypes.byref(nxt),
                                  int(high), flat.numel(), ctypes.c_void_p(flat.data_ptr()))
    _lib.check(rc, 'dccf_confounder_draw')
    words[:] = state
    raw[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0] = left.value
    raw[_OFF_NEXT:_OFF_NEXT + 8].view(np.uint64)[0] = nxt.value
    if generator is not None:
        generator.set_state(blob)
    else:
        torch.set_rng_state(blob)
    return True


def _self_check():
    """Native draw == torch.randint on a private generator: fresh seed (regenerate-first state), mid-generation
    positions, a generation boundary inside the call, and torch continuing correctly from the written-back state."""
    try:
        for seed, high, counts in ((2019, 16000, (7, 1000, 300)), (5, 1, (3,)), (77, (1 << 28) - 1, (700, 11))):
            a, b = torch.Generator(), torch.Generator()
            a.manual_seed(seed)
            b.manual_seed(seed)
            for n in counts:
                want = torch.randint(high, (n,), generator=a)
                got = torch.empty(n, dtype=torch.int64)
                if not _native_into(got, high, b) or not torch.equal(want, got):
                    return False
                if not torch.equal(torch.randint(1 << 20, (5,), generator=a), torch.randint(1 << 20, (5,), generator=b)):
                    return False
                if not torch.equal(torch.randn(3, generator=a), torch.randn(3, generator=b)):
                    return False
        return True
    except Exception:       # noqa: BLE001 — any surprise means: leave the draw to torch
        return False


def available():
    global _checked
    if _checked is None:
        with _lock:
            if _checked is None:
                _checked = _self_check()
    return _checked


def randint(high, shape, out=None, generator=None, min_draws=None):
    """torch.randint(high, shape, out=out) on the CPU generator: identical values, identical generator afterwards."""
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape)) if shape else 1
    if out is None:
        out = torch.empty(shape, dtype=torch.int64)
    threshold = MIN_DRAWS if min_draws is None else min_draws
    usable = (n >= threshold and 0 < high < _MAX_HIGH and out.dtype == torch.int64 and out.device.type == 'cpu'
              and out.is_contiguous() and tuple(out.shape) == shape and available())
    if usable and _native_into(out, high, generator):
        return out
    if generator is not None:
        return torch.randint(high, shape, out=out, generator=generator)
    return torch.randint(high, shape, out=out)


class DeviceStream(object):
    """torch's CPU generator continued on the device for the length of an evaluation pass (dccf_confounder_draw_dev):
    the generator's words are uploaded once, every `draw` is one kernel that writes the ids straight into device memory
    (same numbers torch.randint would have produced on the host, in the same order), and `finish` puts the advanced
    state back into the torch generator.  Between construction and `finish` nothing else may draw from that generator."""

    def __init__(self, device, generator=None):
        self.generator = generator
        self.blob = generator.get_state() if generator is not None else torch.get_rng_state()
        if self.blob.numel() != _STATE_BYTES:
            raise RuntimeError('unexpected torch CPU generator state (%d bytes)' % self.blob.numel())
        raw = self.blob.numpy()
        left = int(raw[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0])
        if not 1 <= left <= _N_WORDS:
            raise RuntimeError('unexpected torch CPU generator state (left = %d)' % left)
        host = np.empty(_N_WORDS + 1, dtype=np.uint32)
        host[:_N_WORDS] = raw[_OFF_WORDS:_OFF_WORDS + 8 * _N_WORDS].view(np.uint64)
        host[_N_WORDS] = _N_WORDS + 1 - left              # index of the next unread word
        self.state = torch.from_numpy(host.view(np.int32)).to(device)
        self.open = True

    def draw(self, high, shape):
        """int64 CUDA tensor == torch.randint(high, shape) on the host generator (0 < high < 2^28)."""
        from . import _lib
        if not self.open:
            raise RuntimeError('DeviceStream used after finish()')
        if not 0 < high < _MAX_HIGH:
            raise ValueError('device confounder draw needs 0 < high < 2^28 (torch draws 64-bit words above)')
        out = torch.empty(tuple(int(x) for x in shape), dtype=torch.int64, device=self.state.device)
        if out.numel():
            lib = _lib.load()
            _lib.check(lib.dccf_confounder_draw_dev(_lib.ptr(self.state), int(high), out.numel(), _lib.ptr(out),
                                                    _lib.stream_ptr()), 'dccf_confounder_draw_dev')
        return out

    def finish(self):
        """Write the advanced state back into the torch generator (one small device->host copy, synchronising)."""
        if not self.open:
            return
        self.open = False
        host = self.state.cpu().numpy().view(np.uint32)
        pos = int(host[_N_WORDS])
        raw = self.blob.numpy()
        raw[_OFF_WORDS:_OFF_WORDS + 8 * _N_WORDS].view(np.uint64)[:] = host[:_N_WORDS]
        raw[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0] = _N_WORDS + 1 - pos
        raw[_OFF_NEXT:_OFF_NEXT + 8].view(np.uint64)[0] = pos
        if self.generator is not None:
            self.generator.set_state(self.blob)
        else:
            torch.set_rng_state(self.blob)
