"""Device-resident batched front-end: ORB extraction + motion-based tracking without leaving HBM.

Thin ctypes layer over plp_orb_extract_batch_dev + plp_tracker_motion_track_batch_dev; used by bench.py and
the pipeline parity tests.  No compute here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import KP_DTYPE, Context, OrbExtractor, PlpError, _P, make_camera, make_grid  # noqa: F401


class TrackLast(C.Structure):
    _fields_ = [("pos_w", _P), ("octave", _P), ("angle", _P), ("desc", _P), ("valid", _P), ("offsets", _P),
                ("pose_pred", _P), ("pose_last", _P)]


class DeviceBuffer:
    """A cudaMalloc'ed block owned through the C ABI (plp_dev_alloc / plp_dev_free)."""

    def __init__(self, ctx: Context, nbytes: int):
        self._ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        ctx._check(ctx._lib.plp_dev_alloc(ctx.handle, C.c_size_t(max(self.nbytes, 1)), C.byref(p)))
        self.ptr = p

    @classmethod
    def from_array(cls, ctx: Context, arr: np.ndarray) -> "DeviceBuffer":
        arr = np.ascontiguousarray(arr)
        buf = cls(ctx, arr.nbytes)
        buf.upload(arr)
        return buf

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self._ctx._check(self._ctx._lib.plp_dev_upload(self._ctx.handle, self.ptr, arr.ctypes.data_as(_P),
                                                       C.c_size_t(arr.nbytes)))

    def download(self, dtype, shape) -> np.ndarray:
        out = np.zeros(shape, dtype)
        assert out.nbytes <= self.nbytes
        self._ctx._check(self._ctx._lib.plp_dev_download(self._ctx.handle, out.ctypes.data_as(_P), self.ptr,
                                                         C.c_size_t(out.nbytes)))
        return out

    def free(self):
        if self.ptr is not None:
            self._ctx._lib.plp_dev_free(self._ctx.handle, self.ptr)
            self.ptr = None


class PinnedBuffer:
    """Page-locked host memory owned through the C ABI (plp_host_alloc_pinned): the source / destination of the
    asynchronous copies of the end-to-end path."""

    def __init__(self, ctx: Context, nbytes: int):
        self._lib = ctx._lib
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        ctx._check(self._lib.plp_host_alloc_pinned(C.c_size_t(max(self.nbytes, 1)), C.byref(p)))
        self.ptr = p

    @classmethod
    def from_array(cls, ctx: Context, arr: np.ndarray) -> "PinnedBuffer":
        arr = np.ascontiguousarray(arr)
        buf = cls(ctx, arr.nbytes)
        C.memmove(buf.ptr, arr.ctypes.data, arr.nbytes)
        return buf

    def view(self, dtype, shape, offset: int = 0) -> np.ndarray:
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        assert offset + n <= self.nbytes
        raw = (C.c_char * n).from_address(self.ptr.value + offset)
        return np.frombuffer(raw, dtype).reshape(shape)

    def free(self):
        if self.ptr is not None:
            self._lib.plp_host_free_pinned(self.ptr)
            self.ptr = None


class FrontEnd:
    """extract (orb_extractor::extract) -> motion_based_track for a batch of frames, all on the device."""

    def __init__(self, ctx: Context, rows: int, cols: int, cam, max_batch: int, max_last_points: int = 4096,
                 max_num_keypts=1000, scale_factor=1.2, num_levels=8, ini_fast_thr=20, min_fast_thr=7,
                 track_ctx: Context | None = None):
        """track_ctx: optional second context (stream) for the tracking kernels; extraction stays on `ctx`.  The two
        streams are chained by plp_ctx_wait_ctx, so extract(k + 1) of ANOTHER FrontEnd can run under track(k)."""
        self.ctx = ctx
        self.track_ctx = track_ctx if track_ctx is not None else ctx
        self.lib = ctx._lib
        self.rows, self.cols, self.max_batch = rows, cols, max_batch
        self.cam = cam
        self.grid = make_grid(cols, rows)
        self.orb = OrbExtractor(ctx, rows, cols, max_num_keypts, scale_factor, num_levels, ini_fast_thr, min_fast_thr,
                                max_batch=max_batch)
        self.cap = self.orb.capacity
        self.max_last = max_last_points
        h = C.c_void_p()
        sf = np.ascontiguousarray(self.orb.scale_factors, np.float32)
        isig = np.ascontiguousarray(self.orb.inv_level_sigma_sq, np.float32)
        ctx._check(self.lib.plp_tracker_create(self.track_ctx.handle, C.byref(cam), C.byref(self.grid), sf.ctypes.data_as(_P),
                                               isig.ctypes.data_as(_P), C.c_int(num_levels), C.c_int(max_batch),
                                               C.c_int(self.cap), C.c_int(max_last_points), C.byref(h)))
        self._trk = h
        B = max_batch
        self.d_imgs = DeviceBuffer(ctx, B * rows * cols)
        self.d_kp = DeviceBuffer(ctx, B * self.cap * KP_DTYPE.itemsize)
        self.d_desc = DeviceBuffer(ctx, B * self.cap * 32)
        self.d_n = DeviceBuffer(ctx, B * 4)
        self.d_status = DeviceBuffer(ctx, B * 4)
        self.d_matched = DeviceBuffer(ctx, B * self.cap * 4)
        self.d_pose = DeviceBuffer(ctx, B * 128)
        self.d_num_valid = DeviceBuffer(ctx, B * 4)
        self.d_n_inl = DeviceBuffer(ctx, B * 4)
        self.d_lm = DeviceBuffer(ctx, B * 4)
        self._last_bufs = []
        self._last = None
        self._last_pinned = []   # host copies of the last-frame arrays (end-to-end path: uploaded every step)
        self._pin_imgs = None
        self._pin_out = None
        self._out_layout = None

    def close(self):
        if self._trk is not None:
            self.lib.plp_tracker_destroy(self._trk)
            self._trk = None
        self.orb.close()

    # -- inputs ---------------------------------------------------------------------------------------
    def upload_images(self, imgs: np.ndarray):
        self.d_imgs.upload(np.ascontiguousarray(imgs, np.uint8))

    def set_last_frames(self, last_list, pose_pred, pose_last):
        """last_list[b]: dict(pos_w[m,3], octave[m], angle[m], desc[m,32], valid[m]|None)."""
        for b in self._last_bufs:
            b.free()
        offs = np.zeros(len(last_list) + 1, np.int32)
        offs[1:] = np.cumsum([len(l["octave"]) for l in last_list])
        assert max(np.diff(offs)) <= self.max_last
        cat = lambda k, dt: np.ascontiguousarray(np.concatenate([np.asarray(l[k], dt) for l in last_list]))
        arrays = [cat("pos_w", np.float64), cat("octave", np.int32), cat("angle", np.float32), cat("desc", np.uint8),
                  np.ascontiguousarray(np.concatenate([np.asarray(l.get("valid") if l.get("valid") is not None else
                                                                  np.ones(len(l["octave"]), np.uint8), np.uint8)
                                                       for l in last_list])),
                  offs, np.ascontiguousarray(pose_pred, np.float64), np.ascontiguousarray(pose_last, np.float64)]
        self._last_bufs = [DeviceBuffer.from_array(self.ctx, a) for a in arrays]
        self._last = TrackLast(*[b.ptr for b in self._last_bufs])
        for b in self._last_pinned:
            b.free()
        self._last_pinned = [PinnedBuffer.from_array(self.ctx, a) for a in arrays]

    # -- end-to-end path: every input of a step comes from pinned host memory, every result goes back ----
    def stage_host_io(self, imgs: np.ndarray):
        """Pinned host staging of one step: the images (input) and one block for all results."""
        batch = imgs.shape[0]
        if self._pin_imgs is not None:
            self._pin_imgs.free()
            self._pin_out.free()
        self._pin_imgs = PinnedBuffer.from_array(self.ctx, np.ascontiguousarray(imgs, np.uint8))
        items = [("kp", self.d_kp, batch * self.cap * KP_DTYPE.itemsize), ("desc", self.d_desc, batch * self.cap * 32),
                 ("n_kp", self.d_n, batch * 4), ("status", self.d_status, batch * 4),
                 ("matched", self.d_matched, batch * self.cap * 4), ("pose", self.d_pose, batch * 128),
                 ("num_valid", self.d_num_valid, batch * 4), ("n_inliers", self.d_n_inl, batch * 4),
                 ("lm_iters", self.d_lm, batch * 4)]
        off, layout = 0, []
        for name, dbuf, nbytes in items:
            layout.append((name, dbuf, off, nbytes))
            off += (nbytes + 255) & ~255
        self._pin_out = PinnedBuffer(self.ctx, off)
        self._out_layout = layout
        self._io_batch = batch

    @property
    def h2d_bytes_per_step(self) -> int:
        return self._pin_imgs.nbytes + sum(b.nbytes for b in self._last_pinned)

    @property
    def d2h_bytes_per_step(self) -> int:
        return sum(nb for _, _, _, nb in self._out_layout)

    def upload_inputs_async(self):
        """H2D of this step's images AND of its last-frame landmarks / descriptors / predicted poses (extraction stream).
        The previous step's tracking kernels and result downloads (tracking stream) must have drained first."""
        lib = self.lib
        if self.track_ctx is not self.ctx:
            self.ctx.wait(self.track_ctx)
        self.ctx._check(lib.plp_dev_upload_async(self.ctx.handle, self.d_imgs.ptr, self._pin_imgs.ptr,
                                                 C.c_size_t(self._pin_imgs.nbytes)))
        for dbuf, pbuf in zip(self._last_bufs, self._last_pinned):
            self.ctx._check(lib.plp_dev_upload_async(self.ctx.handle, dbuf.ptr, pbuf.ptr, C.c_size_t(pbuf.nbytes)))

    def download_outputs_async(self):
        """D2H of everything the host-side data::frame needs from this step: keypoints, descriptors, counts, the
        landmark index kept on every keypoint, pose, valid / inlier counts (tracking stream, after track())."""
        cx = self.track_ctx
        for _, dbuf, off, nbytes in self._out_layout:
            cx._check(self.lib.plp_dev_download_async(cx.handle, C.c_void_p(self._pin_out.ptr.value + off), dbuf.ptr,
                                                      C.c_size_t(nbytes)))

    def host_results(self):
        """Views into the pinned result block (valid after the tracking stream has been synchronised)."""
        b, out = self._io_batch, {}
        shapes = {"kp": (KP_DTYPE, (b, self.cap)), "desc": (np.uint8, (b, self.cap, 32)), "n_kp": (np.int32, (b,)),
                  "status": (np.int32, (b,)), "matched": (np.int32, (b, self.cap)), "pose": (np.float64, (b, 4, 4)),
                  "num_valid": (np.int32, (b,)), "n_inliers": (np.int32, (b,)), "lm_iters": (np.int32, (b,))}
        for name, _, off, _ in self._out_layout:
            dt, shp = shapes[name]
            out[name] = self._pin_out.view(dt, shp, off)
        return out

    # -- the hot path (no host synchronisation) ---------------------------------------------------------
    def extract(self, batch: int):
        if self.track_ctx is not self.ctx:
            self.ctx.wait(self.track_ctx)  # the previous track() still reads the keypoint / descriptor arrays
        self.ctx._check(self.lib.plp_orb_extract_batch_dev(self.orb.handle, self.d_imgs.ptr, C.c_int(batch),
                                                          C.c_size_t(self.cols), self.d_kp.ptr, self.d_desc.ptr,
                                                          self.d_n.ptr, self.d_status.ptr))

    def track(self, batch: int, margin: float = 20.0):
        if self.track_ctx is not self.ctx:
            self.track_ctx.wait(self.ctx)  # the extraction of this batch
        self.ctx._check(self.lib.plp_tracker_motion_track_batch_dev(
            self._trk, C.c_int(batch), self.d_kp.ptr, self.d_desc.ptr, self.d_n.ptr, C.byref(self._last),
            C.c_float(margin), self.d_matched.ptr, self.d_pose.ptr, self.d_num_valid.ptr, self.d_n_inl.ptr,
            self.d_lm.ptr))

    def step(self, batch: int, margin: float = 20.0):
        self.extract(batch)
        self.track(batch, margin)

    # -- results --------------------------------------------------------------------------------------
    def download_keypoints(self, batch: int):
        n = self.d_n.download(np.int32, (batch,))
        kp = self.d_kp.download(KP_DTYPE, (self.max_batch, self.cap))[:batch]
        desc = self.d_desc.download(np.uint8, (self.max_batch, self.cap, 32))[:batch]
        return [(kp[b, :n[b]].copy(), desc[b, :n[b]].copy()) for b in range(batch)]

    def download_tracking(self, batch: int):
        n = self.d_n.download(np.int32, (batch,))
        matched = self.d_matched.download(np.int32, (self.max_batch, self.cap))[:batch]
        return dict(matched=[matched[b, :n[b]].copy() for b in range(batch)],
                    pose=self.d_pose.download(np.float64, (batch, 4, 4)),
                    num_valid=self.d_num_valid.download(np.int32, (batch,)),
                    n_inliers=self.d_n_inl.download(np.int32, (batch,)),
                    lm_iters=self.d_lm.download(np.int32, (batch,)),
                    status=self.d_status.download(np.int32, (batch,)))
