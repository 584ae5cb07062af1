"""CPU: the C-ABI library loads and exports every symbol include/snb200.h declares (no compute calls)."""
import ctypes
import os
import subprocess

from supernormal_b200 import _lib, build


def test_header_symbols_exported():
    path = build.build_library()
    assert os.path.exists(path)
    protos = _lib.parse_header()
    assert len(protos) >= 20 and "snb_march_emit" in protos and "snb_hashgrid_bwd_bwd_input" in protos
    so = ctypes.CDLL(path)
    missing = [n for n in protos if not hasattr(so, n)]
    assert not missing, missing
    exported = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    undeclared = [l.split()[-1] for l in exported.splitlines() if " T snb_" in l and l.split()[-1] not in protos]
    assert not undeclared, f"exported but not in the header: {undeclared}"


def test_host_only_entry_points():
    l = _lib.lib()
    assert l.snb_version() >= 100
    m, total = _lib.make_meta(14, 19, 32, 1.3195079107728942)
    assert total == 5936000 and list(m.resolutions)[:4] == [32, 43, 56, 74]
    import oracle
    s = oracle.hashgrid_spec()
    assert list(m.offsets)[:15] == s.offsets.tolist() and list(m.scales)[:14] == s.scales.tolist()
    # argument errors are reported, not crashed on
    assert l.snb_hashgrid_make_meta(99, 19, 32, 1.5, ctypes.byref(m)) == 0
    assert b"n_levels" in l.snb_last_error()
    assert l.snb_march_count(-1, None, None, None, None, None, 1, 1, 1, None, 0.1, 0.0, None, None) == -2


def test_product_has_no_oracle_import():
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """Every ctypes mirror of a struct in include/snb200.h has the size the C compiler gives it (gcc, same ABI as nvcc's host side)."""
    import ctypes as C
    import subprocess
    from supernormal_b200 import _lib, trainer
    pairs = {"snb_hashgrid_meta": _lib.HashGridMeta, "snb_net": trainer.SnbNet, "snb_patch_batch": trainer.SnbPatchBatch,
             "snb_samples": trainer.SnbSamples, "snb_dataset": trainer.SnbDataset, "snb_batch_out": trainer.SnbBatchOut,
             "snb_train_ctx": trainer.SnbTrainCtx, "snb_peer_group": trainer.SnbPeerGroup}
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "snb200.h"\nint main(void) {\n' +
                   "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in pairs) + "  return 0;\n}\n")
    exe = tmp_path / "sz"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, ct in pairs.items():
        assert int(out[name]) == C.sizeof(ct), (name, out[name], C.sizeof(ct))


def test_sample_buffers_accumulator_sets_and_workspace_size():
    """Host logic of trainer.SampleBuffers (no kernel runs): the two loss-accumulator / totals sets alias one 48-byte allocation each, the
    sample struct of a set points at ITS totals, flip() switches both, and the split backward's workspace has the size the C ABI asks for."""
    import ctypes as C
    import torch
    from supernormal_b200 import _lib
    from supernormal_b200.trainer import P, SampleBuffers
    b = SampleBuffers(n_patches=64, samples_per_ray_cap=16, scratch_stride=128, n_levels=14, device="cpu")
    assert b.capacity == 64 * 16 and b.end_capacity == max(64 * 8, b.capacity // 4)
    need = _lib.lib().snb_sdf_bwd_workspace_bytes(14, b.capacity, b.end_capacity)
    assert b.bwd_ws_bytes == need == P * (b.capacity + b.end_capacity) * (16 + 14 * 8) and b.bwd_ws.numel() == need
    assert _lib.lib().snb_sdf_bwd_workspace_bytes(0, 1, 1) == -1 and _lib.lib().snb_sdf_bwd_workspace_bytes(17, 1, 1) == -1
    for k in (0, 1):
        assert b.set_idx == k
        st = b.stats_totals
        assert st.numel() == 12 and st.dtype == torch.int32
        assert b.stats.data_ptr() == st.data_ptr() and b.stats.dtype == torch.float32 and b.stats.numel() == 8
        assert b.totals.data_ptr() == st.data_ptr() + 32 and b.totals.numel() == 4
        assert b.struct.totals == b.totals.data_ptr() and b.struct.capacity == b.capacity
        b.stats[3] = 2.5
        b.totals[1] = 7
        assert st[3:4].view(torch.float32).item() == 2.5 and st[9].item() == 7
        b.flip()
    assert b.set_idx == 0 and b._sets[0].data_ptr() != b._sets[1].data_ptr()
    assert b.feats.numel() == P * (b.capacity + b.end_capacity) * 16 * 2        # rows of up to 16 half2
