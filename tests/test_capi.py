"""The drop-in boundary without a GPU: libdust_b200.so loads, exports every symbol include/dust_b200.h
declares, the ctypes mirror has the header's struct layouts, host-only entry points work, and there is
no CPU compute fallback (no device => DUST_ERR_NO_DEVICE; the product never touches oracle/)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dust_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dust_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dust_b200 import capi
    lib = capi.load()
    names = _declared_functions()
    assert len(names) >= 17
    assert set(names) == set(capi.SYMBOLS), set(names) ^ set(capi.SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dust_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert lib.dust_abi_version() == 4


def test_ctypes_structs_match_header(tmp_path):
    """sizeof / offsetof as the C compiler sees include/dust_b200.h vs the ctypes mirror"""
    from dust_b200 import capi
    fields = {
        "DustTreeDesc": ["n_levels", "fanout_log2", "root", "root_bytes", "pool_chunks", "n_chunks", "item_bytes", "n_items",
                         "materials", "n_materials", "palette"],
        "DustInstance": ["obj2world", "tree"],
        "DustCamera": ["view_proj", "inverse_view_proj", "camera_view_col0", "position_x", "camera_view_col2", "tan_half_fov", "near"],
        "DustSky": ["r", "g", "b", "direction", "sunlight_intensity"],
        "DustTraceConfig": ["width", "spp", "seed", "shard_index", "flags"],
        "DustFrame": ["radiance", "hit_instance", "hit_leaf", "hit_voxel", "hit_t", "memory", "radiance_format", "gbuffer"],
        "DustGBuffer": ["albedo", "depth", "normal", "motion", "voxel_id"],
        "DustStats": ["primary_rays", "mask_words", "kernel_launches", "mask_words_strict"],
        "DustLaunchInfo": ["kernel", "ms", "primary_rays", "hits_shaded", "mask_words_strict"],
        "DustExposureParams": ["min_log_luminance", "log_luminance_range", "time_coefficient"],
        "DustToneMapParams": ["color_space_conversion", "output_transfer_function", "dst_format"],
        "DustTreeLayout": ["n_levels", "n_nodes16", "n_materials", "off_nodes16", "off_materials", "total_bytes", "bounds_min", "bounds_max"],
    }
    body = "".join(f'printf("{s} %zu\\n", sizeof({s}));\n' + "".join(
        f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));\n' for f in fs) for s, fs in fields.items())
    body += 'printf("DustRay %zu\\nDustHit %zu\\nDustVoxel %zu\\n", sizeof(DustRay), sizeof(DustHit), sizeof(DustVoxel));\n'
    src = tmp_path / "layout.c"
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "dust_b200.h"\nint main(void){{\n{body}return 0;}}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for s, fs in fields.items():
        cls = getattr(capi, s)
        assert int(got[s]) == C.sizeof(cls), s
        for f in fs:
            assert int(got[f"{s}.{f}"]) == getattr(cls, f).offset, (s, f)
    assert got["DustRay"] == "32" and got["DustHit"] == "16" and got["DustVoxel"] == "4"
    assert C.sizeof(capi.DustCamera) == 192 and C.sizeof(capi.DustSky) == 224


def test_shard_layout_is_host_only():
    from dust_b200.pipeline import shard_layout
    assert shard_layout(1920, 1080, 1) == (60 * 34, 60 * 34 * 1024)
    assert shard_layout(1920, 1080, 8) == (255, 255 * 1024)
    assert shard_layout(3840, 2160, 8) == (1020, 1020 * 1024)
    assert shard_layout(33, 1, 2) == (1, 1024)
    from dust_b200 import capi
    assert capi.load().dust_shard_layout(0, 10, 1, None, None) == 1          # DUST_ERR_INVALID_ARGUMENT


def test_no_device_means_error_not_fallback():
    """Without a CUDA device the boundary refuses to create a context; nothing is computed on the CPU."""
    import torch
    from dust_b200 import capi
    lib = capi.load()
    h = C.c_void_p()
    rc = lib.dust_ctx_create(0, C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        lib.dust_ctx_destroy(h)
    else:
        assert rc == 2 and not h.value                                        # DUST_ERR_NO_DEVICE
        from dust_b200.pipeline import Context
        with pytest.raises(capi.DustError):
            Context(0)
    assert lib.dust_ctx_create(0, None) == 1
    assert lib.dust_trace(None, None, None, None, None, None) == 1
    assert lib.dust_last_error_string(None) == b"null context"
    # every entry point added since ABI v2 validates its arguments before touching a device
    assert lib.dust_comm_init(None, None, 0, 1) == 1 and lib.dust_comm_destroy(None) == 1 and lib.dust_comm_unique_id(None) == 1
    assert lib.dust_allgather_framebuffer(None, None, 0, None) == 1
    assert lib.dust_exposure_create(None, None) == 1 and lib.dust_auto_exposure(None, None, None, 0, 0, 0, None, None) == 1
    assert lib.dust_tone_map(None, None, None, None, 0, 0, 0, None, None) == 1
    assert lib.dust_tree_build_voxels(None, None, 0, None, None, 0, None) == 1 and lib.dust_tree_layout(None, None) == 1
    assert lib.dust_scene_update_transforms(None, None, None, 0) == 1


def test_nccl_binding_resolves_at_run_time():
    """dust_comm_unique_id dlopens libnccl.so.2 on first use: with NCCL present it yields 128 non-trivial bytes (no GPU needed),
    without it the call fails cleanly with DUST_ERR_COMM — never a load-time dependency of the library."""
    from dust_b200 import capi
    lib = capi.load()
    buf = (C.c_uint8 * capi.DUST_COMM_ID_BYTES)()
    rc = lib.dust_comm_unique_id(buf)
    assert rc in (0, 6)
    if rc == 0:
        assert any(buf)
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in out


def test_product_never_references_the_oracle():
    """oracle/ is test infrastructure: no source under dust_b200/ or include/ may include, import, link or
    dlopen it (a product path through the oracle would void every parity claim)."""
    bad = []
    for base in ("dust_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".pyc")):
                    continue
                p = os.path.join(dirpath, f)
                text = open(p, errors="ignore").read()
                if re.search(r"pyoracle|liboracle|oracle_(vdb|vox|sky|trace|dense)|from oracle|import oracle|oracle/", text):
                    bad.append(os.path.relpath(p, ROOT))
    assert not bad, bad
    out = subprocess.run(["ldd", os.path.join(ROOT, "dust_b200", "lib", "libdust_host.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
