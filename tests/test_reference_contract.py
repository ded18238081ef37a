"""Static check of the drop-in boundary against the reference's own call sites (runs where
/root/reference is mounted -- the build container; skipped on the GPU box, which has no copy).

The reference's renderer shims build `GaussianRasterizationSettings(...)` and call
`GaussianRasterizer(...)(...)` with keyword arguments and unpack TWO return values
(gs_renderer_single.py:69-95, gs_renderer_multiple.py:95-121); its deformer calls
`lbs_extra(...)` with keywords (sings_hybrid.py:400-406, :526-534).  Parsed with `ast`: every
keyword they pass must be accepted by our classes / functions, and the settings fields they set
must be exactly the required fields of ours."""
import ast
import inspect
import os

import pytest

REF = "/root/reference/sings/rec"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not mounted here")


def _calls(path, name):
    tree = ast.parse(open(path).read())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call):
            f = node.func
            fname = f.id if isinstance(f, ast.Name) else (f.attr if isinstance(f, ast.Attribute) else None)
            if fname == name:
                out.append(node)
    return out


@pytest.mark.parametrize("shim", ["renderer/gs_renderer_single.py", "renderer/gs_renderer_multiple.py"])
def test_renderer_shims_call_our_rasterizer_api(shim):
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    path = os.path.join(REF, shim)
    settings = _calls(path, "GaussianRasterizationSettings")
    assert settings, "the shim constructs the settings"
    required = [f for f in GaussianRasterizationSettings._fields if f not in GaussianRasterizationSettings._field_defaults]
    for c in settings:
        kws = [k.arg for k in c.keywords]
        assert not c.args and sorted(kws) == sorted(required), (kws, required)
    ctor = _calls(path, "GaussianRasterizer")
    assert ctor and all([k.arg for k in c.keywords] == ["raster_settings"] for c in ctor)
    fwd_params = set(inspect.signature(GaussianRasterizer.forward).parameters) - {"self"}
    calls = _calls(path, "rasterizer")
    assert calls
    for c in calls:
        assert not c.args and {k.arg for k in c.keywords} <= fwd_params, [k.arg for k in c.keywords]
    # the result is unpacked into exactly two names: (rendered_image, radii)
    tree = ast.parse(open(path).read())
    unpacked = [n for n in ast.walk(tree) if isinstance(n, ast.Assign) and isinstance(n.value, ast.Call)
                and getattr(n.value.func, "id", None) == "rasterizer"]
    assert unpacked and all(isinstance(n.targets[0], ast.Tuple) and len(n.targets[0].elts) == 2 for n in unpacked)


def test_deformer_call_sites_match_lbs_extra_signature():
    from sings_b200.deform import lbs_extra
    ours = inspect.signature(lbs_extra).parameters
    calls = _calls(os.path.join(REF, "models/sings_hybrid.py"), "lbs_extra")
    assert len(calls) >= 2
    for c in calls:
        assert len(c.args) <= 2                                    # (A, v_shaped) positionally
        assert {k.arg for k in c.keywords} <= set(ours), [k.arg for k in c.keywords]
    # and the reference's own definition has the same parameter names in the same order
    ref_def = [n for n in ast.walk(ast.parse(open(os.path.join(REF, "utils/body_model/lbs.py")).read()))
               if isinstance(n, ast.FunctionDef) and n.name == "lbs_extra"][0]
    assert [a.arg for a in ref_def.args.args] == list(ours)
