"""Frame-index selection (host side, pure integer arithmetic, bit-exact with the reference).

The reference samples *decoded frame lists*; here the same rules produce INDEX lists so that the
gather can be fused into the GPU ingest kernel (b2_ingest_u8 `frame_index`), with -1 meaning an
all-zero frame.  Reference: medsos_lrcn/src/loader_data.py:35-51 (uniform_sampling /
duplicate_frames; same code lrcn/ucf50-lrcn.py:84-100), lrcn/backup_ucf50.py:52-62 (seek variant),
lrcn/lrcn.py:149-155 (crime variant with zero-frame padding)."""
from typing import List, Optional


def uniform_sampling(n_frames: int, sequence_length: int) -> List[int]:
    """Indices kept by `uniform_sampling(frames, T)`: all frames when n <= T, else every
    (n // T)-th frame, first T of them."""
    if n_frames <= sequence_length:
        return list(range(n_frames))
    interval = n_frames // sequence_length
    return list(range(0, n_frames, interval))[:sequence_length]


def duplicate_frames(indices: List[int], sequence_length: int) -> List[int]:
    """`duplicate_frames`: cycle a short clip until it has T frames."""
    indices = list(indices)
    if len(indices) >= sequence_length:
        return indices[:sequence_length]
    if not indices:
        raise ValueError("cannot extend an empty clip (the reference loops forever here)")
    reps = -(-sequence_length // len(indices))
    return (indices * reps)[:sequence_length]


def medsos_indices(n_frames: int, sequence_length: int) -> List[int]:
    """loader_data.py:171-178: uniform sampling, then cycling when the clip is short."""
    idx = uniform_sampling(n_frames, sequence_length)
    if len(idx) < sequence_length:
        idx = duplicate_frames(idx, sequence_length)
    return idx


def seek_indices(n_frames: int, sequence_length: int) -> Optional[List[int]]:
    """UCF50 seek sampling (backup_ucf50.py:52-62): None when the clip is skipped (n < T)."""
    if n_frames < sequence_length:
        return None
    interval = n_frames // sequence_length
    return [i * interval for i in range(sequence_length)]


def crime_indices(n_frames: int, sequence_length: int) -> List[int]:
    """lrcn/lrcn.py:151-155: strided selection, short clips padded with zero frames (-1)."""
    if n_frames >= sequence_length:
        interval = n_frames // sequence_length
        return list(range(0, n_frames, interval))[:sequence_length]
    return list(range(n_frames)) + [-1] * (sequence_length - n_frames)
