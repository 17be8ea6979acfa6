"""ctypes loader of the C-ABI library (libses3d.so, include/ses3d.h). No fallback: if the
library is missing or no CUDA device is present, the calls fail loudly."""
import ctypes as C
import os
from pathlib import Path

from .layouts import AssocDump, Params, PriorParams, SynthConfig

PKG = Path(__file__).resolve().parent
# SES3D_LIB: development override (A/B builds of the kernels, scripts/build_variants.py); the product path is fixed
LIB_PATH = Path(os.environ["SES3D_LIB"]) if os.environ.get("SES3D_LIB") else PKG / "libses3d.so"

EXPORTS = ("ses3d_default_params", "ses3d_create", "ses3d_destroy", "ses3d_get_tables", "ses3d_triangulate_batch",
           "ses3d_reproject_batch", "ses3d_process_batch", "ses3d_process_batch_ragged", "ses3d_reserve", "ses3d_munkres_batch", "ses3d_launch_count",
           "ses3d_set_profiling", "ses3d_last_kernel_ms", "ses3d_last_error_string", "ses3d_version",
           "ses3d_synth_frames", "ses3d_synth_frames_device", "ses3d_assembler_default_config",
           "ses3d_assembler_create", "ses3d_assembler_destroy", "ses3d_assembler_add", "ses3d_assembler_pop",
           "ses3d_assembler_stats", "ses3d_mailbox_replay", "ses3d_wire_decode_person2dlist", "ses3d_wire_encode_person2dlist",
           "ses3d_wire_decode_personcovlist", "ses3d_wire_encode_personcovlist", "ses3d_prior_default_params",
           "ses3d_prior_create", "ses3d_prior_destroy", "ses3d_prior_reset", "ses3d_prior_run", "ses3d_prior_get_tracks",
           "ses3d_prior_launch_count", "ses3d_prior_last_kernel_ms", "ses3d_measure_fma_peak", "ses3d_markers_batch", "ses3d_overlay_batch", "ses3d_prior_run_ragged",
           "ses3d_n_cams", "ses3d_check", "ses3d_bind_thread_to_device_numa", "ses3d_create_multi", "ses3d_multi_destroy",
           "ses3d_multi_device_count", "ses3d_multi_handle", "ses3d_multi_process_batch",
           "ses3d_multi_process_batch_ragged")


class Ses3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ses3d error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libses3d.so (build it first with `python -m smartedgesensor3dhumanpose_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(f"{LIB_PATH} not built: run `python -m smartedgesensor3dhumanpose_b200.build` "
                                "(the CUDA library is required, there is no CPU path)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, u32, i64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_int64
    L.ses3d_default_params.argtypes = [C.POINTER(Params)]
    L.ses3d_default_params.restype = None
    L.ses3d_create.argtypes = [i32, vp, C.POINTER(Params), i32, C.POINTER(vp)]
    L.ses3d_destroy.argtypes = [vp]
    L.ses3d_get_tables.argtypes = [vp, vp, vp]
    L.ses3d_triangulate_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, C.POINTER(AssocDump), u32, vp]
    L.ses3d_reproject_batch.argtypes = [vp, i32, i32, vp, vp, vp, vp, u32, vp]
    L.ses3d_process_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, C.POINTER(AssocDump), u32, vp]
    L.ses3d_process_batch_ragged.argtypes = [vp, i32, i32, vp, vp, i32, vp, i64, vp, vp, i64, vp, C.POINTER(i64),
                                             C.POINTER(i64), u32]
    L.ses3d_reserve.argtypes = [vp, i32, i32, i32]
    L.ses3d_munkres_batch.argtypes = [vp, i32, i32, i32, vp, vp]
    L.ses3d_launch_count.argtypes = [vp]
    L.ses3d_launch_count.restype = i64
    L.ses3d_set_profiling.argtypes = [vp, i32]
    L.ses3d_last_kernel_ms.argtypes = [vp, vp]
    L.ses3d_last_error_string.restype = C.c_char_p
    L.ses3d_version.restype = C.c_char_p
    L.ses3d_synth_frames.argtypes = [i32, vp, C.POINTER(SynthConfig), i64, i32, vp, vp, vp, vp]
    L.ses3d_synth_frames_device.argtypes = [i32, vp, C.POINTER(SynthConfig), i64, i32, vp, vp, vp, vp]
    L.ses3d_assembler_default_config.argtypes = [i32, vp]
    L.ses3d_assembler_create.argtypes = [vp, C.POINTER(vp)]
    L.ses3d_assembler_destroy.argtypes = [vp]
    L.ses3d_assembler_add.argtypes = [vp, i32, i64, i64]
    L.ses3d_assembler_pop.argtypes = [vp, vp, vp, vp, vp]
    L.ses3d_assembler_stats.argtypes = [vp, vp]
    L.ses3d_mailbox_replay.argtypes = [i32, vp, vp, vp, vp]
    sz = C.c_size_t
    L.ses3d_wire_decode_person2dlist.argtypes = [vp, sz, vp, vp, vp, sz, vp, vp, i32]
    L.ses3d_wire_encode_person2dlist.argtypes = [u32, i64, C.c_char_p, C.c_float, vp, i32, vp, sz]
    L.ses3d_wire_encode_person2dlist.restype = sz
    L.ses3d_wire_decode_personcovlist.argtypes = [vp, sz, vp, vp, vp, sz, vp, vp, i32, vp, vp, i32]
    L.ses3d_wire_encode_personcovlist.argtypes = [u32, i64, C.c_char_p, i32, vp, vp, vp, i32, vp, sz]
    L.ses3d_wire_encode_personcovlist.restype = sz
    L.ses3d_prior_default_params.argtypes = [C.POINTER(PriorParams)]
    L.ses3d_prior_default_params.restype = None
    L.ses3d_prior_create.argtypes = [C.POINTER(PriorParams), i32, i32, i32, C.POINTER(vp)]
    L.ses3d_prior_destroy.argtypes = [vp]
    L.ses3d_prior_reset.argtypes = [vp]
    L.ses3d_prior_run.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, u32, vp]
    L.ses3d_prior_get_tracks.argtypes = [vp, i32, vp, vp]
    L.ses3d_prior_run_ragged.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, i64, vp, vp, C.POINTER(i64)]
    L.ses3d_prior_launch_count.argtypes = [vp]
    L.ses3d_prior_launch_count.restype = i64
    L.ses3d_prior_last_kernel_ms.argtypes = [vp, vp]
    L.ses3d_measure_fma_peak.argtypes = [i32, i32, C.POINTER(C.c_double)]
    L.ses3d_markers_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, u32, vp]
    L.ses3d_overlay_batch.argtypes = [vp, i32, i32, vp, vp, i32, i32, vp, u32, vp]
    L.ses3d_n_cams.argtypes = [vp]
    L.ses3d_check.argtypes = [vp]
    L.ses3d_bind_thread_to_device_numa.argtypes = [i32, C.POINTER(i32)]
    L.ses3d_create_multi.argtypes = [i32, vp, C.POINTER(Params), i32, vp, C.POINTER(vp)]
    L.ses3d_multi_destroy.argtypes = [vp]
    L.ses3d_multi_device_count.argtypes = [vp]
    L.ses3d_multi_handle.argtypes = [vp, i32]
    L.ses3d_multi_handle.restype = vp
    L.ses3d_multi_process_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, C.POINTER(AssocDump)]
    L.ses3d_multi_process_batch_ragged.argtypes = [vp, i32, i32, vp, vp, i32, vp, i64, vp, vp, i64, vp, vp, vp]
    _lib = L
    return L


def measure_fma_peak(device=0, fp64=False):
    """Measured FMA-pipe peak of the GPU in TFLOP/s (micro-benchmark inside libses3d.so)."""
    v = C.c_double(0)
    check(load().ses3d_measure_fma_peak(device, int(fp64), C.byref(v)))
    return v.value


def bind_thread_to_device_numa(device=0):
    """Pin the calling thread to the CPUs of the GPU's NUMA node (no-op without sysfs); returns the node or -1."""
    node = C.c_int32(-1)
    check(load().ses3d_bind_thread_to_device_numa(device, C.byref(node)))
    return node.value


def check(rc):
    if rc != 0:
        raise Ses3dError(rc, load().ses3d_last_error_string().decode())
