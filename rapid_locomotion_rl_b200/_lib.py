"""ctypes binding of include/rl_b200.h.

The struct layouts are GENERATED from the header text at import time (single source of
truth) and checked against `rl_sizeof()` of the loaded library.  There is no fallback: if
librl_b200.so is missing and cannot be built, importing a product path raises.
"""
import ctypes as C
import os
import re

from . import build as _build

_HEADER = os.path.join(_build.INCLUDE, "rl_b200.h")

_SCALARS = {
    "int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "uint8_t": C.c_uint8,
    "float": C.c_float, "double": C.c_double, "int": C.c_int, "uint16_t": C.c_uint16, "int16_t": C.c_int16,
    "int8_t": C.c_int8,
}


def _parse_header(path):
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(\w+)\s+\(?(-?\d+)\)?\s*$", text, flags=re.M)}
    enums = {}
    for m in re.finditer(r"enum\s+\w+\s*\{(.*?)\};", text, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = [x.strip() for x in item.split("=")]
                nxt = int(val)
            else:
                name = item
            enums[name] = nxt
            nxt += 1
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            decl = decl.replace("const ", "")
            base, rest = decl.split(" ", 1)
            base_ptr = base.endswith("*")
            base = base.rstrip("*")
            for d in rest.split(","):
                d = d.strip()
                is_ptr = base_ptr or d.startswith("*")
                d = d.lstrip("* ")
                am = re.match(r"(\w+)\[(\w+)\]$", d)
                elem = C.c_void_p if is_ptr else (structs[base] if base in structs else _SCALARS[base])
                if am:
                    name = am.group(1)
                    dim = am.group(2)
                    n = int(dim) if dim.isdigit() else defines[dim]
                    ctype = elem * n
                else:
                    name = d
                    ctype = elem
                fields.append((name, ctype))
        structs[m.group(3)] = type(m.group(3), (C.Structure,), {"_fields_": fields})
    return defines, enums, structs


DEFINES, ENUMS, STRUCTS = _parse_header(_HEADER)
RlEnvCfg = STRUCTS["RlEnvCfg"]
RlEnvBuffers = STRUCTS["RlEnvBuffers"]
RlResetCfg = STRUCTS["RlResetCfg"]
RlResetBuffers = STRUCTS["RlResetBuffers"]
RlGacCfg = STRUCTS["RlGacCfg"]
RlGacBuffers = STRUCTS["RlGacBuffers"]
RlWgradProblem = STRUCTS["RlWgradProblem"]
RlStorageAdd = STRUCTS["RlStorageAdd"]
RlPeerComm = STRUCTS["RlPeerComm"]
RlChainTensor = STRUCTS["RlChainTensor"]
RlChainLoadOp = STRUCTS["RlChainLoadOp"]
RlChainMmaOp = STRUCTS["RlChainMmaOp"]
RlChainEpiOp = STRUCTS["RlChainEpiOp"]
RlChainDesc = STRUCTS["RlChainDesc"]
RlChainPpoLoss = STRUCTS["RlChainPpoLoss"]
RlRolloutBoundary = STRUCTS["RlRolloutBoundary"]
RlRolloutAct = STRUCTS["RlRolloutAct"]

RL_OK = DEFINES["RL_OK"]
REWARD_TERM_IDS = {k[len("RL_REW_"):].lower(): v for k, v in ENUMS.items() if k.startswith("RL_REW_") and k != "RL_REW_COUNT"}

# every exported symbol of the header: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "rl_last_error": (C.c_char_p, []),
    "rl_version": (C.c_char_p, []),
    "rl_sizeof": (C.c_int64, [C.c_char_p]),
    "rl_debug_env_trace": (C.c_int, [C.c_int32, _P]),
    "rl_debug_env_rows": (C.c_int, [C.c_int32]),
    "rl_debug_env_rows_trace": (C.c_int, [C.c_int32, _P, C.c_int32]),
    "rl_env_torques": (C.c_int, [_P, _P, _P]),
    "rl_env_post_physics": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "rl_env_step_fused": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "rl_env_heights": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, _P, _P]),
    "rl_env_reset": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "rl_gac_scatter": (C.c_int, [_P, _P, _P]),
    "rl_gac_update_sample": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "rl_gae_workspace_bytes": (C.c_int64, [C.c_int32]),
    "rl_gae_scan": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P, _P]),
    "rl_gae_normalize": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    "rl_gae": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P]),
    "rl_storage_add": (C.c_int, [_P, _P]),
    "rl_rollout_boundary": (C.c_int, [_P, _P]),
    "rl_rollout_act": (C.c_int, [_P, _P]),
    "rl_history_push": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "rl_gemm_bf16": (C.c_int, [_P, _P, _P, _P, _P, _P] + [C.c_int32] * 10 + [_P]),
    "rl_ppo_gather": (C.c_int, [_P] * 11 + [C.c_int32] * 4 + [_P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, _P]),
    "rl_ppo_gather_history": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P]),
    "rl_cast_bf16": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "rl_ppo_loss": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, C.c_int32, C.c_float, C.c_float, C.c_float,
                              C.c_int32, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "rl_adapt_loss": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P, _P]),
    "rl_grad_finalize": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_double, C.c_float, C.c_float, C.c_int32, _P, _P, _P]),
    "rl_adam": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32,
                          C.c_float, _P, _P]),
    "rl_gemm_init": (C.c_int, []),
    "rl_grad_finalize_from_norm": (C.c_int, [_P, _P, _P, C.c_double, C.c_float, C.c_float, C.c_int32, _P, _P, _P]),
    "rl_enable_peer_access": (C.c_int, [C.c_int32]),
    "rl_peer_allreduce": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, _P, _P, C.c_uint32, _P]),
    "rl_wgrad_grouped": (C.c_int, [_P, C.c_int32, _P]),
    "rl_chain_create": (C.c_int, [_P, _P]),
    "rl_chain_run": (C.c_int, [_P, C.c_int32, _P]),
    "rl_chain_run_tiles": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "rl_chain_destroy": (C.c_int, [_P]),
    "rl_chain_set_ppo_loss": (C.c_int, [_P, _P]),
    "rl_chain_trace": (C.c_int, [_P, C.c_int32]),
    "rl_chain_read_trace": (C.c_int64, [_P, _P, C.c_int64]),
    "rl_adam_shadows": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32,
                                  C.c_float, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "rl_refresh_shadows": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "rl_policy_sample": (C.c_int, [_P, _P, C.c_int32, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P, _P]),
}


def declared_symbols():
    """Function names declared in the header (used by the CPU test that the .so exports them all)."""
    text = re.sub(r"/\*.*?\*/", "", open(_HEADER).read(), flags=re.S)
    text = re.sub(r"typedef\s+struct.*?\}\s*\w+\s*;", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rl_\w+)\s*\(", text)))


class RlError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (building first if the sources changed) the native library.  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not _build.is_current():
        try:
            path = _build.build()
        except Exception as exc:  # no nvcc on this box: use the shipped .so if there is one
            if not os.path.exists(_build.LIB):
                raise RlError("librl_b200.so is missing and could not be built: %s" % exc)
            path = _build.LIB
    handle = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    for sname, st in STRUCTS.items():
        n = handle.rl_sizeof(sname.encode())
        if n != C.sizeof(st):
            raise RlError("layout mismatch for %s: header binding %d B, library %d B" % (sname, C.sizeof(st), n))
    _lib = handle
    return _lib


def check(rc):
    if rc != RL_OK:
        raise RlError("rl_b200 error %d: %s" % (rc, lib().rl_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
