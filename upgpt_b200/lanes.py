"""Lanes: several independent batches in flight on one GPU.

At the benchmarked batch size the denoising step is a chain of ~300 dependent launches, most of them bound by latency, not by the
machine: a second, independent chain fills the first one's bubbles (two B = 8 U-Net steps side by side: 6.40 ms instead of 2 x 4.02 ms,
bit-identical results; three: 9.15 ms -- profiles/r02_two_batches_in_flight.json). A lane is one such chain:

    with lanes.lane(1):                      # everything enqueued here goes to lane 1's stream
        z, _ = DDIMSampler(model).sample(...)
        img = model.decode_first_stage(z)

Lane 0 is the caller's current stream (the default: nothing changes for single-stream users). Lane i > 0 runs on the library's
auxiliary stream 2 i and forks its parallel branches (a ResBlock's skip GEMM) onto auxiliary stream 2 i + 1, so every stream of every
lane has its own scratch slot in the library (common.cuh: stream_slot). Engines -- activation buffers, recorded programs, captured
graphs, cond-cache -- exist per lane (the engine cache key carries the lane); the packed weights are shared by all lanes.
Single host thread: the lanes are fed one after the other, asynchronously; nothing here synchronises with the device.
"""
import contextlib
import ctypes as C

MAX_LANES = 8
THROUGHPUT_SM_WEIGHT = 8.0
_current = 0


def set_throughput_mode(on=True, sm_weight=None):
    """Library settings for several batches in flight: the GEMM tiler counts SM time (upgpt_gemm_set_sm_weight), so that a small layer
    takes the SMs it needs instead of all it can use, and launches go without programmatic dependent launch (upgpt_set_pdl).
    Call it before engines are built, and do not switch it while engines built under the other setting are still in use: captured
    graphs keep the tiling they were built with, but an EAGER replay of a recorded program re-plans its GEMMs, and the folded
    LayerNorm's row-statistic slots (= N tiles of the producing GEMM) were sized at record time. Measured on the bbox.yaml path, B = 8 (profiles/r02_throughput_mode.txt): 4 lanes 65.9 images/s
    (one batch in flight: 31) vs the latency objective's 49.0 with 3 lanes (38.9 with one)."""
    import os
    from . import _C
    w = (THROUGHPUT_SM_WEIGHT if sm_weight is None else float(sm_weight)) if on else 0.0
    _C.check(_C.lib().upgpt_gemm_set_sm_weight(C.c_double(w)), "upgpt_gemm_set_sm_weight")
    # programmatic dependent launch makes a successor's CTAs resident while the predecessor drains: they wait on SMs that another lane's
    # kernel could use (63.8 -> 64.9 images/s without it, 4 lanes; profiles/r02_knobs_under_lanes.txt). One batch alone wants it on.
    if os.environ.get("UPGPT_PDL") is None:
        _C.check(_C.lib().upgpt_set_pdl(0 if on else 1), "upgpt_set_pdl")
    return w


def current():
    return _current


def branch_aux(lane_id):
    """Index of the library's auxiliary stream that carries lane `lane_id`'s forked branches."""
    return 0 if lane_id == 0 else 2 * lane_id + 1


def stream(lane_id):
    """torch stream of a lane (lane 0: the current stream)."""
    import torch
    if lane_id == 0:
        return torch.cuda.current_stream()
    from . import _C
    if not 0 < lane_id < MAX_LANES:
        raise ValueError("lane %d out of range (0 .. %d)" % (lane_id, MAX_LANES - 1))
    lib = _C.lib()
    lib.upgpt_aux_stream.restype = C.c_void_p
    ptr = lib.upgpt_aux_stream(2 * lane_id)
    if not ptr:
        raise _C.UpgptError("no auxiliary stream for lane %d" % lane_id)
    return torch.cuda.ExternalStream(ptr)


@contextlib.contextmanager
def lane(lane_id):
    """Makes `lane_id` the current lane and its stream torch's current stream."""
    import torch
    global _current
    prev = _current
    s = stream(lane_id)
    _current = lane_id
    try:
        with torch.cuda.stream(s):
            yield s
    finally:
        _current = prev
