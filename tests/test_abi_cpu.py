"""CPU-side checks of the drop-in boundary: libvvae.so loads, exports every symbol include/vvae.h declares, the ctypes
structs mirror the header, and the product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vvae.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vvae_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    names = _declared()
    assert len(names) >= 40
    for must in ("vvae_gemm", "vvae_attn_fwd", "vvae_attn_bwd", "vvae_conv3d_fwd", "vvae_conv3d_dgrad", "vvae_conv3d_wgrad",
                 "vvae_layernorm_fwd", "vvae_layernorm_bwd", "vvae_qknorm_rope_fwd", "vvae_groupnorm_silu_fwd",
                 "vvae_reparam_gate_fwd", "vvae_recon_loss_fwd", "vvae_kl_fwd", "vvae_adam_step", "vvae_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from video_vae_b200 import _ffi
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in include/vvae.h but not exported by libvvae.so: {missing}"
    # and the Python binding binds exactly the declared set
    assert sorted(_ffi.EXPORTED) == _declared()


def _struct_fields(name):
    src = open(HEADER).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    return fields


@pytest.mark.parametrize("cname,pyname", [("vvae_gemm_args", "GemmArgs"), ("vvae_attn_args", "AttnArgs"),
                                          ("vvae_conv_args", "ConvArgs")])
def test_ctypes_structs_mirror_the_header(cname, pyname):
    from video_vae_b200 import _ffi
    assert [f[0] for f in getattr(_ffi, pyname)._fields_] == _struct_fields(cname)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import video_vae_b200 as V
    from video_vae_b200 import _ffi
    assert _ffi.lib.vvae_device_ok() == 0
    assert _ffi.lib.vvae_version() > 0
    with pytest.raises(_ffi.VvaeError):
        _ffi.require_device()
    # building the module tree works on CPU (parameters only) but any forward must refuse to run
    m = V.VideoVAE(32, 32, 3, 16, 1, 1, 64, 2, 32, 8, 8, 4, V.Rngs(0), dtype=torch.float32)
    x = torch.zeros(1, 2, 32, 32, 3)
    mask = torch.ones(1, 2, dtype=torch.bool)
    with pytest.raises(_ffi.VvaeError):
        m(x, mask[:, None, None, :], V.Rngs(0), train=False)


def test_integration_md_names_exist():
    """Every replacement named in INTEGRATION.md's mapping table (`video_vae_b200.mod.{a,b}` / `video_vae_b200.mod.name`)
    imports: the drop-in guide cannot drift from the package."""
    import importlib
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    found = 0
    for mod, names in re.findall(r"`video_vae_b200\.([a-z_]+)\.\{([A-Za-z0-9_,]+)\}`", text):
        m = importlib.import_module("video_vae_b200." + mod)
        for n in names.split(","):
            assert hasattr(m, n), f"video_vae_b200.{mod}.{n}"
            found += 1
    for mod, name in re.findall(r"`video_vae_b200\.([a-z_]+)\.([A-Za-z_][A-Za-z0-9_]*)`", text):
        m = importlib.import_module("video_vae_b200." + mod)
        assert hasattr(m, name), f"video_vae_b200.{mod}.{name}"
        found += 1
    assert found >= 30
    import video_vae_b200 as V
    for n in ("VideoVAE", "Rngs", "loss_fn", "DEFAULT_HPARAMS"):       # the top-level names the guide's snippet imports
        assert hasattr(V, n)


# ---------------------------------------------------------------------------------------------- vvae_comm_* (csrc/comm.cu)
def test_comm_token_and_argument_checks_without_gpu():
    """The gradient-exchange entry points: NCCL resolves through dlopen (no link-time dependency), rank
    0's rendezvous token can be made on the host, and every call that would touch a device or a bogus handle fails with a
    status and a message instead of crashing."""
    from video_vae_b200 import _ffi
    lib = _ffi.lib
    tok = (ctypes.c_char * 128)()
    assert lib.vvae_comm_unique_id(ctypes.cast(tok, ctypes.c_void_p)) == 0, lib.vvae_last_error()
    tok2 = (ctypes.c_char * 128)()
    assert lib.vvae_comm_unique_id(ctypes.cast(tok2, ctypes.c_void_p)) == 0
    assert bytes(tok) != bytes(128) and bytes(tok) != bytes(tok2)          # a real, fresh token each time
    assert lib.vvae_comm_unique_id(None) == 1                              # VVAE_ERR_INVALID
    h = ctypes.c_void_p()
    assert lib.vvae_comm_init(ctypes.byref(h), ctypes.cast(tok, ctypes.c_void_p), 2, 2) == 1
    assert b"rank 2 of 2" in lib.vvae_last_error()
    if not torch.cuda.is_available():
        assert lib.vvae_comm_init(ctypes.byref(h), ctypes.cast(tok, ctypes.c_void_p), 0, 1) == 2   # VVAE_ERR_CUDA
        assert h.value is None
    assert lib.vvae_comm_allreduce(None, None, 0, 0, 1, None) == 1
    assert lib.vvae_comm_broadcast(None, None, 0, 0, 0, None) == 1
    assert b"not a communicator" in lib.vvae_last_error()
    assert lib.vvae_comm_destroy(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_native_comm_refuses_without_gpu_and_file_exchange(tmp_path):
    import threading
    from video_vae_b200 import _ffi, ddp
    with pytest.raises(_ffi.VvaeError):
        ddp.NativeComm(0, 1)
    # the token transport of NativeComm.from_env: rank 0 writes atomically, the others poll
    path = str(tmp_path / "token")
    token = bytes(range(128))
    got = {}
    t = threading.Thread(target=lambda: got.setdefault("r1", ddp.NativeComm.file_exchange(path, 10.0)(None)))
    t.start()
    assert ddp.NativeComm.file_exchange(path)(token) == token
    t.join(20)
    assert got["r1"] == token
    with pytest.raises(TimeoutError):
        ddp.NativeComm.file_exchange(str(tmp_path / "absent"), 0.2)(None)



def test_workspace_bytes_front_door():
    """vvae_workspace_bytes(op, dims, ndims) agrees with the per-operator size functions and rejects bad input."""
    from video_vae_b200 import _ffi
    lib = _ffi.lib

    def ws(op, *dims):
        arr = (ctypes.c_longlong * max(len(dims), 1))(*dims)
        return lib.vvae_workspace_bytes(op, arr, len(dims))
    assert ws(0, 128, 64, 64, 128, 64) == lib.vvae_convT122_workspace_bytes(128, 64, 64, 128, 64) > 0
    assert ws(1, 170_630_000) == 4 * lib.vvae_sumsq_partials(170_630_000) > 0
    assert ws(2, 128, 8, 256) == 4 * 128 * 8 * 256
    assert ws(2, 128, 8) == -1 and ws(0, 1, 2, 3, 4, 0) == -1 and ws(9, 1) == -1
    assert lib.vvae_workspace_bytes(1, None, 1) == -1
