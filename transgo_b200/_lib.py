"""ctypes binding of libtransgo_b200.so (the C ABI declared in include/transgo_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtransgo_b200.so")

_lib = None

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p


class TransgoLibraryError(RuntimeError):
    pass


def _declare(L):
    L.tg_version.restype = ctypes.c_int
    L.tg_last_error.restype = ctypes.c_char_p
    L.tg_supported_board.argtypes = [ctypes.c_int]
    L.tg_state_bytes.argtypes = [ctypes.c_int]
    L.tg_legal_words.argtypes = [ctypes.c_int]
    L.tg_set_board_size.argtypes = [ctypes.c_int]
    L.tg_set_board_size.restype = ctypes.c_bool
    L.tg_env_reset.argtypes = [ctypes.c_int, c_vp, c_i64, c_vp]
    L.tg_env_step.argtypes = [ctypes.c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp]
    L.tg_env_atari_words.argtypes = [ctypes.c_int]
    L.tg_env_step_inc.argtypes = [ctypes.c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, c_vp]
    L.tg_env_legal.argtypes = [ctypes.c_int, c_vp, c_i64, c_vp, ctypes.c_int, ctypes.c_int, c_vp]
    L.tg_env_encode.argtypes = [ctypes.c_int, c_vp, c_i64, c_vp, ctypes.c_int, c_vp]
    L.tg_env_score.argtypes = [ctypes.c_int, c_vp, c_i64, ctypes.c_float, c_vp, c_vp, ctypes.c_int, c_vp]
    L.tg_env_playout.argtypes = [ctypes.c_int, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, c_i64,
                                 ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp]
    L.tg_env_playout_scored.argtypes = [ctypes.c_int, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, c_i64,
                                        ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_float, ctypes.c_int, c_vp]
    L.tg_env_playout_io.argtypes = [ctypes.c_int, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, c_i64, ctypes.c_int,
                                    c_vp, c_vp, c_vp, ctypes.c_float, ctypes.c_int, c_vp]
    L.tg_env_playout_tile.argtypes = [ctypes.c_int, c_i64, ctypes.c_int]
    L.tg_sub_encode.argtypes = [ctypes.c_int, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]
    L.tg_traj_assemble.argtypes = [ctypes.c_int, ctypes.c_int, c_i64, c_i64] + [c_vp] * 11 + [c_i64, c_i64, ctypes.c_int, c_vp, c_vp]
    L.tg_dirichlet.argtypes = [c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint32, c_i64, c_vp]
    L.tg_net_add_bn_relu.argtypes = [c_vp, c_vp, ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_vp]
    L.tg_net_add_bn_relu_f32.argtypes = [c_vp, c_vp, ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_vp]
    L.tg_net_attention.argtypes = [c_vp, c_vp, ctypes.c_float, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]
    L.tg_net_attention_fused2.argtypes = [c_vp, c_vp, c_vp, ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]
    L.tg_net_stem.argtypes = [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_vp]
    L.tg_net_conv3x3_head.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]
    L.tg_net_attention2.argtypes = [c_vp, c_vp, ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]
    L.tg_net_attention_fused.argtypes = [c_vp, c_vp, c_vp, ctypes.c_float, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]
    # the reference's 15 symbols (go_env.h:23-71), typed exactly as GoEnv/environment.py:43-90 types them
    L.Init.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]
    L.Reset.argtypes = [c_vp]
    L.Step.argtypes = [c_vp, c_vp, ctypes.c_int]
    L.Step.restype = ctypes.c_bool
    L.Step_.argtypes = [c_vp, ctypes.c_int]
    L.Step_.restype = ctypes.c_bool
    L.checkAction.argtypes = [c_vp, ctypes.c_int]
    L.checkAction.restype = ctypes.c_bool
    L.isTerminated.argtypes = [c_vp]
    L.isTerminated.restype = ctypes.c_bool
    L.Encode.argtypes = [c_vp, c_vp]
    L.Encode.restype = ctypes.c_bool
    L.getScore.argtypes = [c_vp]
    L.getScore.restype = ctypes.c_float
    L.getTerritory.argtypes = [c_vp, c_vp]
    L.getTerritory.restype = ctypes.c_float
    L.getLegalAction.argtypes = [c_vp, c_vp]
    L.getLegalAction.restype = ctypes.c_int
    L.getLegalNoEye.argtypes = [c_vp, c_vp]
    L.getLegalNoEye.restype = ctypes.c_int
    L.Show.argtypes = [c_vp]
    L.getPlayer.argtypes = [c_vp]
    L.getPlayer.restype = ctypes.c_int      # environment.py:82 reads the uint8_t as c_int
    L.getStep.argtypes = [c_vp]
    L.getStep.restype = ctypes.c_int
    L.getSubEncode.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.getSubEncode.restype = None


def lib():
    """Load the CUDA library or fail loudly -- there is no CPU path to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TransgoLibraryError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C transgo_b200/csrc`). transgo_b200 has no CPU fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(code, what=""):
    if code != 0:
        raise TransgoLibraryError("%s failed (%d): %s" % (what or "libtransgo_b200 call", code,
                                                          lib().tg_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None, device=None):
    """cudaStream_t of torch's current stream.  The C library launches on the CURRENT CUDA device (it never calls
    cudaSetDevice), so an object whose tensors live on `device` must be driven with that device current: anything else
    would run kernels on one GPU against another GPU's pointers.  Fail loudly instead."""
    import torch
    if device is not None:
        dev = torch.device(device)
        idx = torch.cuda.current_device() if dev.index is None else dev.index
        if idx != torch.cuda.current_device():
            raise TransgoLibraryError("tensors live on cuda:%d but the current device is cuda:%d: wrap the call in "
                                      "`with torch.cuda.device(%d)` (or torch.cuda.set_device)" % (idx, torch.cuda.current_device(), idx))
        s = torch.cuda.current_stream(idx) if stream is None else stream
    else:
        s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)
