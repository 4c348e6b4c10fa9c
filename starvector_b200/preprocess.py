"""GPU image preprocessing behind the reference's processor interface.

`ImageTrainProcessor` keeps the constructor and `__call__` of the reference's class of the same name (reference
starvector/data/util.py:40-53: `ImageTrainProcessor(mean=None, std=None, size=224)`, `processor(pil_image) -> [3,S,S]`)
and `SimpleStarVectorProcessor` those of starvector_arch.py:16-73 (`processor(images=...) -> {"pixel_values": ...}`);
both run on the C-ABI `sv_preproc_*` entry points: the uint8 HWC bytes PIL holds are uploaded and the alpha paste / pad /
bicubic resize / ToTensor / Normalize happen in two CUDA kernels, bit-identical to Pillow + torchvision.  Results are
DEVICE tensors.  There is no CPU implementation here: without the CUDA library or a GPU the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)     # data/util.py:33-36
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _as_u8_hwc(item) -> np.ndarray:
    """PIL image (modes RGB / RGBA, what the reference's transforms handle) or uint8 ndarray [H,W,3|4] -> ndarray view."""
    if isinstance(item, np.ndarray):
        arr = item
    elif isinstance(item, torch.Tensor):
        arr = item.detach().cpu().numpy()
    elif hasattr(item, "mode") and hasattr(item, "size"):
        if item.mode not in ("RGB", "RGBA"):
            raise ValueError(f"image mode {item.mode!r}: the reference's transform yields 3 channels only for RGB/RGBA input; "
                             "convert the image first")
        arr = np.asarray(item)
    else:
        raise ValueError(f"unsupported image type {type(item).__name__}: pass a PIL image or a uint8 [H,W,3|4] array")
    if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] not in (3, 4):
        raise ValueError(f"images must be uint8 [H,W,3|4]; got {arr.dtype} {tuple(arr.shape)}")
    if arr.strides[2] != 1 or arr.strides[1] != arr.shape[2] or arr.strides[0] < arr.shape[1] * arr.shape[2]:
        arr = np.ascontiguousarray(arr)
    return arr


class GpuImageProcessor:
    """One `sv_preproc` handle: `run(list of images) -> [n,3,S,S]` on the device."""

    def __init__(self, size: int = 224, mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None,
                 alpha: str = "white", pad_square: bool = True, device: int = 0, dtype: torch.dtype = torch.float32):
        if alpha not in ("white", "drop"):
            raise ValueError("alpha must be 'white' (ImageTrainProcessor) or 'drop' (SimpleStarVectorProcessor)")
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be torch.float32 or torch.bfloat16")
        self.size, self.dtype = int(size), dtype
        self.mean = tuple(CLIP_MEAN if mean is None else mean)
        self.std = tuple(CLIP_STD if std is None else std)
        self.device = torch.device("cuda", device)
        self._lib = _lib.load()
        self._lock = threading.Lock()
        desc = _lib.PreprocDesc(self.size, _lib.SV_ALPHA_WHITE if alpha == "white" else _lib.SV_ALPHA_DROP, int(bool(pad_square)),
                                _lib.SV_DTYPE_F32 if dtype == torch.float32 else _lib.SV_DTYPE_BF16,
                                (C.c_float * 3)(*self.mean), (C.c_float * 3)(*self.std))
        h = C.c_void_p()
        self._h = None
        self._ck(self._lib.sv_preproc_create(C.byref(desc), device, C.byref(h)))
        self._h = h

    def _ck(self, code: int) -> None:
        if code == _lib.SV_OK:
            return
        msg = self._lib.sv_preproc_last_error(self._h)
        text = msg.decode("utf-8", "replace") if msg else ""
        if code == _lib.SV_ERR_INVALID:
            raise ValueError(f"starvector_b200: {text}")
        raise _lib.EngineError(f"starvector_b200 (code {code}): {text}")

    def run(self, images: Iterable) -> torch.Tensor:
        arrays = [_as_u8_hwc(im) for im in images]
        n = len(arrays)
        if n == 0:
            raise ValueError("no images")
        descs = (_lib.ImageU8 * n)()
        for i, a in enumerate(arrays):
            tight = a.shape[1] * a.shape[2]
            descs[i] = _lib.ImageU8(a.ctypes.data, a.shape[1], a.shape[0], a.shape[2], 0 if a.strides[0] == tight else a.strides[0])
        out = torch.empty((n, 3, self.size, self.size), dtype=self.dtype, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        with self._lock:
            if self._h is None:
                raise _lib.EngineError("processor is closed")
            self._ck(self._lib.sv_preproc_run_host(self._h, descs, n, C.c_void_p(out.data_ptr()), C.c_void_p(stream.cuda_stream)))
            stream.synchronize()          # the host arrays were borrowed for (possibly pageable) async copies: keep them alive until done
        return out

    def launch_count(self) -> int:
        return int(self._lib.sv_preproc_launch_count(self._h)) if self._h is not None else 0

    def close(self) -> None:
        with self._lock:
            if self._h is not None:
                self._lib.sv_preproc_destroy(self._h)
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ImageTrainProcessor(GpuImageProcessor):
    """reference starvector/data/util.py:40-66: RGBA pasted on white, pad to square (255), PIL bicubic resize, ToTensor,
    Normalize.  `processor(img)` -> device tensor [3,size,size]; `processor.batch(imgs)` -> [n,3,size,size]."""

    def __init__(self, mean=None, std=None, size: int = 224, device: int = 0, dtype: torch.dtype = torch.float32, **kwargs):
        super().__init__(size=size, mean=mean, std=std, alpha="white", pad_square=True, device=device, dtype=dtype)

    def __call__(self, item) -> torch.Tensor:
        return self.run([item])[0]

    def batch(self, items: Iterable) -> torch.Tensor:
        return self.run(items)


class SimpleStarVectorProcessor(GpuImageProcessor):
    """reference starvector/model/starvector_arch.py:16-90 (image half): RGBA -> `convert("RGB")`, otherwise the same
    transform; `processor(images=...)` returns `{"pixel_values": tensor}` ([3,S,S] for one image, [n,3,S,S] for a list).
    Text is the tokenizer's business (`text=` is accepted only as None)."""

    def __init__(self, tokenizer=None, size: int = 224, mean=None, std=None, device: int = 0, dtype: torch.dtype = torch.float32, **kwargs):
        super().__init__(size=size, mean=mean, std=std, alpha="drop", pad_square=True, device=device, dtype=dtype)
        self.tokenizer = tokenizer

    def __call__(self, images=None, text=None, max_length=None, **kwargs) -> dict:
        if images is None and text is None:
            raise ValueError("You have to specify at least one of `images` or `text`.")     # starvector_arch.py:62-63
        if text is not None:
            raise NotImplementedError("text inputs go through the tokenizer; only `images=` is processed here")
        if isinstance(images, (list, tuple)):
            return {"pixel_values": self.run(images)}
        return {"pixel_values": self.run([images])[0]}


class _Features(dict):
    """`BatchFeature` stand-in: a dict whose keys are also attributes (`out.pixel_values`)."""

    __getattr__ = dict.__getitem__


class SiglipImageProcessor(GpuImageProcessor):
    """The image half of `AutoProcessor.from_pretrained("google/siglip-*")` the reference builds for v2 (reference
    starvector/model/image_encoder/image_encoder.py:32-48, used at :119): `convert("RGB")`, PIL bicubic resize straight to
    (size,size) without padding, rescale 1/255, normalise with mean = std = 0.5.  Bit-identical to the PIL-based
    `SiglipImageProcessorPil` of the installed transformers (the only implementation in the pinned 4.49); the
    torchvision-backed default of transformers 5.x resamples in floating point and differs by up to 2/255 per value.
    `processor(images=..., return_tensors="pt").pixel_values` -> [n,3,size,size] on the device."""

    def __init__(self, size: int = 384, image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), device: int = 0,
                 dtype: torch.dtype = torch.float32, **kwargs):
        if isinstance(size, dict):
            if size.get("height") != size.get("width"):
                raise ValueError("only square output sizes are built")
            size = size["height"]
        super().__init__(size=size, mean=image_mean, std=image_std, alpha="drop", pad_square=False, device=device, dtype=dtype)

    def __call__(self, images=None, return_tensors: Optional[str] = "pt", **kwargs) -> _Features:
        if images is None:
            raise ValueError("You have to specify `images`.")
        if not isinstance(images, (list, tuple)):
            images = [images]
        return _Features(pixel_values=self.run(images))


ImageLike = Union[np.ndarray, "torch.Tensor", object]
__all__: List[str] = ["GpuImageProcessor", "ImageTrainProcessor", "SimpleStarVectorProcessor", "SiglipImageProcessor", "CLIP_MEAN", "CLIP_STD"]
