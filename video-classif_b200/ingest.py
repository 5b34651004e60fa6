"""GPU-side replacement of the reference's per-frame host preprocessing
(cv2.resize -> BGR2RGB -> /255 -> float32 -> permute; loader_data.py:162-163,182,201,112)."""
from typing import Sequence

import torch

from . import ops, sampling


def ingest_clip(frames_u8: torch.Tensor, sequence_length: int, height: int, width: int, mode: str = "medsos",
                out_dtype=torch.float32, divisor: float = 255.0) -> torch.Tensor:
    """One decoded video (uint8 [n,H0,W0,3] BGR, on the GPU) -> float [T,3,height,width].

    mode 'medsos' (RGB, cycle short clips), 'crime' (keeps BGR, zero-pads short clips),
    'seek' (UCF50: skips short clips -> ValueError)."""
    n = frames_u8.shape[0]
    if mode == "medsos":
        idx, swap = sampling.medsos_indices(n, sequence_length), True
    elif mode == "crime":
        idx, swap = sampling.crime_indices(n, sequence_length), False
    elif mode == "seek":
        idx, swap = sampling.seek_indices(n, sequence_length), True
        if idx is None:
            raise ValueError(f"clip with {n} frames is skipped by the seek sampler (needs >= {sequence_length})")
    else:
        raise ValueError(mode)
    index = torch.tensor(idx, dtype=torch.int32, device=frames_u8.device)
    return ops.ingest_u8(frames_u8, height, width, frame_index=index, out_dtype=out_dtype, swap_rb=swap, divisor=divisor)


def ingest_batch(clips_u8: torch.Tensor, height: int, width: int, out_dtype=torch.float32, swap_rb: bool = True,
                 divisor: float = 255.0, out=None) -> torch.Tensor:
    """Already-sampled uint8 clips [B,T,H0,W0,3] -> [B,T,3,height,width] in one launch.  out: write into this tensor (e.g.
    model.encoder_input_buffer(...): the encoder graph's static input, no copy between ingest and the encoder pass)."""
    B, T = clips_u8.shape[:2]
    out = ops.ingest_u8(clips_u8.reshape(B * T, *clips_u8.shape[2:]), height, width, out_dtype=out_dtype,
                        swap_rb=swap_rb, divisor=divisor, out=out)
    return out.reshape(B, T, 3, height, width)
