"""The C-ABI shared library: loads without a GPU, exports every symbol include/dmn_b200.h declares, and its
host-only entry points (plan construction, parameter table, argument validation) behave."""
import ctypes as C
import os
import re

import pytest
import torch

from diffusion_model_nemo_b200 import _lib as L
from conftest import CFGS, ROOT
from oracle import ref_port as O


def header_functions():
    src = open(os.path.join(ROOT, "include", "dmn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dmn_b200.h but not exported"
    assert sorted(L.SYMBOLS) == names, "ctypes binding table and header disagree"
    assert lib.dmn_abi_version() == 1


def _cfg(cfg, size, batch, act=L.ACT_F32, eng=L.CONV_SIMT):
    c = L.UnetCfg()
    c.dim, c.n_mults = cfg["dim"], len(cfg["dim_mults"])
    for i, m in enumerate(cfg["dim_mults"]):
        c.dim_mults[i] = m
    c.channels = cfg["channels"]
    c.out_dim = cfg["channels"] * (2 if cfg.get("learned_variance") else 1)
    c.groups, c.with_time_emb = cfg["groups"], 1
    c.num_classes = -1 if cfg.get("num_classes") is None else cfg["num_classes"]
    c.image_size, c.max_batch, c.act_dtype, c.conv_engine, c.max_time_rows = size, batch, act, eng, 1000
    return c


@pytest.mark.parametrize("name", list(CFGS))
@pytest.mark.parametrize("mode", [(L.ACT_F32, L.CONV_SIMT), (L.ACT_BF16, L.CONV_TCGEN05)])
def test_plan_parameter_table_matches_reference_state_dict(name, mode):
    cfg, size, b = CFGS[name]
    lib = L.lib()
    h = C.c_void_p()
    L.check(lib.dmn_plan_create(C.byref(_cfg(cfg, size, b, *mode)), C.byref(h)))
    try:
        got = {}
        for i in range(lib.dmn_plan_num_params(h)):
            shp = (C.c_int64 * 4)()
            nd = lib.dmn_plan_param_shape(h, i, C.byref(shp))
            got[lib.dmn_plan_param_name(h, i).decode()] = tuple(shp[k] for k in range(nd))
        assert got == {k: tuple(v) for k, v in O.unet_param_shapes(cfg).items()}
        assert lib.dmn_plan_weights_bytes(h) > 4 * sum(int(torch.tensor(v).prod()) for v in got.values()) * (0.45 if mode[0] else 0.99)
        assert lib.dmn_plan_workspace_bytes(h) > 0
        assert lib.dmn_plan_ready(h) == 0          # nothing bound / loaded yet
    finally:
        lib.dmn_plan_destroy(h)


def test_argument_validation_is_loud():
    lib = L.lib()
    h = C.c_void_p()
    bad = _cfg(CFGS["tiny"][0], 16, 2)
    bad.groups = 5
    assert lib.dmn_plan_create(C.byref(bad), C.byref(h)) == -1
    assert b"groups" in lib.dmn_last_error()
    bad = _cfg(CFGS["tiny"][0], 16, 2, L.ACT_F32, L.CONV_TCGEN05)
    assert lib.dmn_plan_create(C.byref(bad), C.byref(h)) == -1
    bad = _cfg(CFGS["cfg1"][0], 30, 2)             # 30 -> 15 -> odd
    assert lib.dmn_plan_create(C.byref(bad), C.byref(h)) == -1
    with pytest.raises(ValueError):
        L.check(-1, "x")
    with pytest.raises(NotImplementedError):
        L.check(-2, "x")
    # forward before bind/load is a state error, not a crash
    good = _cfg(CFGS["tiny"][0], 16, 2)
    L.check(lib.dmn_plan_create(C.byref(good), C.byref(h)))
    assert lib.dmn_unet_forward(h, None, None, None, None, 1, None) == -4
    lib.dmn_plan_destroy(h)


@pytest.mark.parametrize("mode", [(L.ACT_F32, L.CONV_SIMT), (L.ACT_BF16, L.CONV_TCGEN05)])
def test_film_plan_parameter_table_and_layout(mode):
    """WaveGradUNet plans (cfg.film = 1): the engine's parameter table is the subset of the reference state_dict it evaluates -- no
    time MLP, FiLM layers 0 .. n_levels-1 (the bottleneck FiLM is discarded by the reference and the last n_levels-1 are never called,
    unet.py:204-210,247) -- and dmn_plan_film_layout reports their channels in table-column order."""
    from conftest import WG_CFGS

    cfg, size, b = WG_CFGS["wg_cfg"]
    lib = L.lib()
    c = _cfg(cfg, size, b, *mode)
    c.with_time_emb, c.film = 0, 1
    h = C.c_void_p()
    L.check(lib.dmn_plan_create(C.byref(c), C.byref(h)))
    try:
        names = {lib.dmn_plan_param_name(h, i).decode() for i in range(lib.dmn_plan_num_params(h))}
        ref = set(O.unet_param_shapes(cfg))
        assert names <= ref and not any(n.startswith("time_mlp") or ".mlp." in n for n in names)
        films = sorted({int(n.split(".")[1]) for n in names if n.startswith("films.")})
        assert films == [0, 1, 2, 3]
        assert (ref - names) == {n for n in ref if n.startswith("films.") and int(n.split(".")[1]) >= 4}
        ch = (C.c_int32 * 16)()
        n = lib.dmn_plan_film_layout(h, ch, 16)
        assert [ch[i] for i in range(n)] == [128, 128, 256, 256]
        kinds = []
        for i in range(lib.dmn_plan_num_ops(h)):
            k, e, fl, by = C.c_int32(), C.c_int32(), C.c_double(), C.c_double()
            L.check(lib.dmn_plan_op_info(h, i, None, 0, C.byref(k), C.byref(e), C.byref(fl), C.byref(by)))
            kinds.append(k.value)
        assert kinds.count(7) == 4                                   # one modulation per evaluated FiLM layer
    finally:
        lib.dmn_plan_destroy(h)
    # a FiLM plan must not carry a time embedding (unet.py:195)
    c.with_time_emb = 1
    assert lib.dmn_plan_create(C.byref(c), C.byref(h)) == -1
