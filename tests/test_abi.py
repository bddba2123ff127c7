"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/dmme_b200.h declares; argument errors come back as codes + messages, not crashes; the
Python classes keep the reference's constructor signatures and state_dict layout."""
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dmme_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmme_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dmme_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dmme_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.dmme_abi_version() == 8


def test_argument_errors_are_reported_not_fatal():
    import ctypes as C
    from dmme_b200 import _lib
    lib = _lib.load()
    assert lib.dmme_conv2d_fwd(None, None) == -1
    assert b"null descriptor" in lib.dmme_last_error()
    rc = lib.dmme_groupnorm_fwd(None, None, 8, 0, 1, 4, 2, 1e-5, None, None, None, None, 0, 0, None, 1, None, 1, None, None, None)
    assert rc == -1 and b"groupnorm" in lib.dmme_last_error()
    d = _lib.ConvDesc()
    d.kernel = 99
    assert lib.dmme_conv2d_fwd(C.byref(d), None) == -1
    with pytest.raises(RuntimeError, match="conv2d_fwd"):
        _lib.check(-1, "conv2d_fwd")


def test_conv_desc_struct_matches_header_field_order():
    from dmme_b200 import _lib
    text = open(os.path.join(ROOT, "include", "dmme_b200.h")).read()
    body = re.search(r"typedef struct dmme_conv_desc \{(.*?)\} dmme_conv_desc;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(void|float|int|long long|dmme_out_norm|dmme_sampler_epilogue)\s*\*?\s*", "", decl)
        names += [re.sub(r"\[\d+\]$", "", n.strip().lstrip("*").strip()) for n in decl.split(",")]
    assert names == [f[0] for f in _lib.ConvDesc._fields_]
    # the nested dmme_out_norm mirrors its header declaration too, and both structs have the C compiler's size
    body = re.search(r"typedef struct dmme_out_norm \{(.*?)\} dmme_out_norm;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            decl = re.sub(r"^(const\s+)?(void|float|int|long long)\s*\*?\s*", "", decl)
            names += [n.strip().lstrip("*").strip() for n in decl.split(",")]
    assert names == [f[0] for f in _lib.OutNorm._fields_]
    import ctypes
    assert ctypes.sizeof(_lib.OutNorm) == 64 and ctypes.sizeof(_lib.ConvDesc) == 200 + 2 * 64 + 8
    assert ctypes.sizeof(_lib.SamplerEpilogue) == 88


def test_cpu_tensors_are_refused_loudly():
    from dmme_b200.models.ddpm import UNet
    m = UNet(pos_dim=4, emb_dim=8, num_groups=2, channels_per_depth=(4, 8), num_blocks=1).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 3, 8, 8), torch.tensor([1]))


def test_constructor_signatures_match_reference_defaults():
    from dmme_b200 import DDIM, DDPM, IDDPM
    from dmme_b200.models import ddpm, iddpm

    def defaults(fn):
        return {k: v.default for k, v in inspect.signature(fn).parameters.items() if k != "self"}

    d = defaults(ddpm.UNet.__init__)
    assert list(d)[:8] == ["in_channels", "pos_dim", "emb_dim", "num_groups", "dropout", "channels_per_depth",
                           "num_blocks", "attention_depths"]
    assert (d["in_channels"], d["pos_dim"], d["emb_dim"], d["num_groups"], d["dropout"]) == (3, 128, 512, 32, 0.1)
    assert d["channels_per_depth"] == (128, 256, 256, 256) and d["num_blocks"] == 2 and d["attention_depths"] == (2,)
    di = defaults(iddpm.UNet.__init__)
    assert di["dropout"] == 0.3 and di["attention_depths"] == (2, 3)
    assert defaults(DDPM.__init__) == {"model": inspect._empty, "timesteps": 1000, "start": 0.0001, "end": 0.02}
    assert defaults(DDIM.__init__) == {"model": inspect._empty, "timesteps": 1000, "sub_timesteps": 50,
                                       "tau_schedule": "quadratic"}
    assert defaults(IDDPM.__init__) == {"model": inspect._empty, "timesteps": 1000, "loss_type": "hybrid", "gamma": 0.001,
                                        "schedule": "cosine", "offset": 0.008, "start": 0.0001, "end": 0.02}
    with pytest.raises(NotImplementedError):
        DDIM(torch.nn.Identity(), tau_schedule="cubic")
    with pytest.raises(NotImplementedError):
        IDDPM(torch.nn.Identity(), schedule="sigmoid")


def test_schedule_buffers_are_not_serialised_and_state_dict_prefix():
    from dmme_b200 import DDIM
    from dmme_b200.models.ddpm import UNet
    d = DDIM(UNet(pos_dim=4, emb_dim=8, num_groups=2, channels_per_depth=(4, 8), num_blocks=1))
    keys = list(d.state_dict().keys())
    assert all(k.startswith("model.") for k in keys)
    assert "model.condition.0.embeddings" in keys
    assert d.beta.shape == (1001, 1, 1, 1) and d.tau.dtype == torch.int64
