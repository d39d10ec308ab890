"""Host-buffer serving loop: double-buffered input staging around the regressor.

The reference demo (regressor/demo.py) moves every batch to the GPU, runs the model and reads the result back
in sequence.  Here the three stages run on three CUDA streams: while batch i is in the network, batch i + 1 is
copied host->device from pinned memory and the results of batch i - 1 are copied device->host, so a stream of
batches costs max(copy, compute) per batch instead of their sum.  Results are bit-identical to calling the
model directly (same kernels, same order per batch).
"""
from typing import Callable, Dict, Iterable, List, Optional

import torch

MEAS_KEYS = ('mass', 'height', 'chest', 'waist', 'hips')


def pack_result(o) -> Dict[str, torch.Tensor]:
    """The tensors a caller reads back per batch: final-stage vertices and betas, and the 5 measurements."""
    st = o[o['stage_keys'][-1]]          # the reference's own list of stage entries (iterative_regressor.py)
    return dict(vertices=st['vertices'], betas=st['betas'],
                measurements=torch.stack([o['measurements'][k] for k in MEAS_KEYS], 1))


class HostPipeline:
    """model: a shapy_b200 regressor on `device`; depth input slots (2 = double buffering)."""

    def __init__(self, model, device, depth: int = 2, input_stage=None):
        if torch.device(device).type != 'cuda':
            raise RuntimeError('HostPipeline needs a CUDA device (there is no CPU path)')
        self.model, self.device, self.depth = model, torch.device(device), depth
        self.input_stage = input_stage     # shapy_b200.preprocess.InputStage: enables submit_u8()
        self.u8_slots = [None] * depth
        self.desc_slots = [None] * depth
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.slots: List[Optional[torch.Tensor]] = [None] * depth
        self.in_ready = [torch.cuda.Event() for _ in range(depth)]
        self.slot_free = [torch.cuda.Event() for _ in range(depth)]
        self.n = 0
        self.last_out = torch.cuda.Event()

    def _new_slot(self, shape, dtype, cur):
        """Device staging buffer of the copy-in stream.  The caching allocator hands out memory that is free in the
        ALLOCATING stream's order: a block released by Python while the compute stream still has kernels pending on
        it (an intermediate of the previous batch) is legal for a new compute-stream tensor, but not for a buffer that
        the copy-in stream writes right away.  (Found the hard way: the descriptor table of batch 1 landed in a block
        batch 0's head kernel was still writing, and the crop kernel then read garbage offsets.)  So the buffer is
        allocated under the copy-in stream, that stream first catches up with the compute stream, and the buffer's use
        by the compute stream is recorded so that a later free waits for it."""
        self.s_in.wait_stream(cur)
        with torch.cuda.stream(self.s_in):
            t = torch.empty(shape, dtype=dtype, device=self.device)
        t.record_stream(cur)
        return t

    def submit(self, images_host: torch.Tensor, out_host: Dict[str, torch.Tensor],
               between: Optional[Callable[[], None]] = None):
        """Enqueues one batch: pinned host images in, pinned host result tensors out (filled asynchronously; call
        drain() before reading them).  `between` runs on the compute stream before the forward (bench: L2 flush)."""
        if not images_host.is_pinned():
            raise ValueError('HostPipeline.submit: images must be in pinned host memory')
        cur = torch.cuda.current_stream(self.device)
        slot = self.n % self.depth
        if self.slots[slot] is None or self.slots[slot].shape != images_host.shape:
            self.slots[slot] = self._new_slot(images_host.shape, images_host.dtype, cur)
        with torch.cuda.stream(self.s_in):
            if self.n >= self.depth:
                self.s_in.wait_event(self.slot_free[slot])        # the forward that last read this slot is done
            self.slots[slot].copy_(images_host, non_blocking=True)
            self.in_ready[slot].record(self.s_in)
        cur.wait_event(self.in_ready[slot])
        if between is not None:
            between()
        with torch.no_grad():
            res = pack_result(self.model(self.slots[slot]))
        self.slot_free[slot].record(cur)
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            for k, t in res.items():
                t.record_stream(self.s_out)
                out_host[k].copy_(t, non_blocking=True)
            self.last_out.record(self.s_out)
        self.n += 1

    def submit_u8(self, images_u8_host: torch.Tensor, desc_host: torch.Tensor, out_host: Dict[str, torch.Tensor],
                  between: Optional[Callable[[], None]] = None):
        """uint8 entry: pinned host uint8 image bytes + the pinned crop-descriptor table of
        `InputStage` (uniform_table / pack) cross PCIe (4x fewer bytes than fp32 crops); crop, resize and normalisation
        run on the device (`shapy_preprocess_forward`) on the compute stream in front of the forward."""
        if self.input_stage is None:
            raise RuntimeError('HostPipeline.submit_u8 needs an InputStage (input_stage=...)')
        if not (images_u8_host.is_pinned() and desc_host.is_pinned()):
            raise ValueError('HostPipeline.submit_u8: images and descriptors must be in pinned host memory')
        import ctypes as C
        from . import _lib
        n = desc_host.numel() // C.sizeof(_lib.ImageDesc)
        cur = torch.cuda.current_stream(self.device)
        slot = self.n % self.depth
        if self.u8_slots[slot] is None or self.u8_slots[slot].numel() != images_u8_host.numel():
            self.u8_slots[slot] = self._new_slot((images_u8_host.numel(),), torch.uint8, cur)
        if self.desc_slots[slot] is None or self.desc_slots[slot].numel() != desc_host.numel():
            self.desc_slots[slot] = self._new_slot((desc_host.numel(),), torch.uint8, cur)
        with torch.cuda.stream(self.s_in):
            if self.n >= self.depth:
                self.s_in.wait_event(self.slot_free[slot])
            self.u8_slots[slot].copy_(images_u8_host.reshape(-1), non_blocking=True)
            self.desc_slots[slot].copy_(desc_host, non_blocking=True)
            self.in_ready[slot].record(self.s_in)
        cur.wait_event(self.in_ready[slot])
        if between is not None:
            between()
        with torch.no_grad():
            x = self.input_stage.run_device(self.u8_slots[slot], self.desc_slots[slot], n)
            res = pack_result(self.model(x))
        self.slot_free[slot].record(cur)
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            for k, t in res.items():
                t.record_stream(self.s_out)
                out_host[k].copy_(t, non_blocking=True)
            self.last_out.record(self.s_out)
        self.n += 1

    def drain(self):
        """Makes the current stream wait for every outstanding device->host copy (then synchronise to read)."""
        torch.cuda.current_stream(self.device).wait_event(self.last_out)


def run_stream(model, device, batches: Iterable[torch.Tensor], outs: List[Dict[str, torch.Tensor]], depth: int = 2):
    """Convenience wrapper: runs all pinned host `batches` through the pipeline into the pinned `outs`."""
    pipe = HostPipeline(model, device, depth)
    for x, o in zip(batches, outs):
        pipe.submit(x, o)
    pipe.drain()
    torch.cuda.synchronize(device)
    return outs
