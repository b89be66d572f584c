"""In-process driver for an ACE-emitted model unit on the B200 runtime.

`ace_compiler_b200/models/lib<model>.so` is the reference's checked-in, unmodified
`<model>.onnx.inc` compiled against this repo's header tree (tests/build_emitted.py); it links
libace_b200.so.  Loading it with RTLD_GLOBAL lets the runtime find the unit's callbacks
(Get_context_params, Main_graph, ...) exactly as the linker does for the stand-alone
executables.  The calls below are the reference's own driver API (fhe-cmplr/rtlib/include/
common/rt_api.h:24-68): Prepare_context, Prepare_input, Run_main_graph, Handle_output,
Finalize_context.  One model per process, like the reference's global context."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def model_path(model):
    return os.path.join(_HERE, "models", "lib%s.so" % model)


def synthetic_image(idx=0):
    """LCG x = x*1664525 + 1013904223, seed 12345 + idx, -> [-0.5, 0.5)"""
    x = (12345 + idx) & 0xFFFFFFFF
    out = np.zeros(3 * 32 * 32)
    for i in range(out.size):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = float(x >> 8) / 16777216.0 - 0.5
    return out


class EmittedModel:
    def __init__(self, model, data_file, device=0, own_keys=True, seed=None, even_poly=True,
                 quiet=True):
        path = model_path(model)
        if not os.path.exists(path):
            raise RuntimeError("%s is not built (run python __graft_entry__.py where the "
                               "reference tree is available)" % path)
        os.environ["ACE_B200_DATA_FILE"] = data_file
        os.environ["ACE_B200_NO_KEYGEN"] = "0" if own_keys else "1"
        if seed is not None:
            os.environ["ACE_B200_SEED"] = str(seed)
        os.environ.setdefault("RTLIB_BTS_EVEN_POLY", "1" if even_poly else "0")
        os.environ["ACE_B200_QUIET"] = "1" if quiet else "0"
        self.lib = L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.Ace_set_device.argtypes = [C.c_int]
        L.Alloc_tensor.restype = C.c_void_p
        L.Alloc_tensor.argtypes = [C.c_size_t] * 4 + [C.c_void_p]
        L.Free_tensor.argtypes = [C.c_void_p]
        L.Prepare_input.argtypes = [C.c_void_p, C.c_char_p]
        L.Handle_output.restype = C.POINTER(C.c_double)
        L.Handle_output.argtypes = [C.c_char_p]
        L.Ace_timer_stop_ms.restype = C.c_float
        L.Ace_launch_count.restype = C.c_uint64
        L.Ace_context.restype = C.c_void_p
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        L.Ace_set_device(device)
        L.Prepare_context()

    def prepare_input(self, image, name="input"):
        """encode + public-key encrypt on the GPU (Prepare_input, rtlib.c:41-54)"""
        v = np.ascontiguousarray(image, dtype=np.float64)
        t = self.lib.Alloc_tensor(1, 3, 32, 32, v.ctypes.data)
        self.lib.Prepare_input(t, name.encode())
        self.lib.Free_tensor(t)

    def run(self):
        self.lib.Run_main_graph()

    def handle_output(self, n, name="output"):
        """decrypt + decode (Handle_output, rtlib.c:56-72)"""
        p = self.lib.Handle_output(name.encode())
        out = np.array([p[i] for i in range(n)])
        self.libc.free(p)
        return out

    def timer_start(self):
        self.lib.Ace_timer_start()

    def timer_stop_ms(self):
        return float(self.lib.Ace_timer_stop_ms())

    def launch_count(self):
        return int(self.lib.Ace_launch_count())

    def close(self):
        self.lib.Finalize_context()
