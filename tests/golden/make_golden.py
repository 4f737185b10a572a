"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the authoring container only:

    make -C oracle ref && python tests/golden/make_golden.py

Outputs (tests/golden/):
    sample_query.txt / sample_subject.txt   the reference's own fixture (original/BGSA_CPU/sample-data)
    golden_semiglobal_knc.npz               the reference's only checked-in result
                                            (banded/BGSA_KNC/data/result.txt + .info), reordered to [query][subject]
    ref_vectors.npz                         inputs + outputs of the reference builds on small seeded sets
"""
import shutil
import struct
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent / "tools"))
import refutil as R  # noqa: E402
import synth  # noqa: E402

REF = Path("/root/reference")


def main():
    shutil.copy(REF / "original/BGSA_CPU/sample-data/query.txt", HERE / "sample_query.txt")
    shutil.copy(REF / "original/BGSA_CPU/sample-data/subject.txt", HERE / "sample_subject.txt")
    # --- the checked-in golden: 1 block, 2 devices, 3 queries, 64+64 subjects
    raw = np.fromfile(REF / "banded/BGSA_KNC/data/result.txt", dtype=np.int16)
    info = (REF / "banded/BGSA_KNC/data/result.txt.info").read_bytes()
    nblocks, ndev, nq = struct.unpack("<iiq", info[:16])
    c0, c1, extra = struct.unpack("<qqi", info[16:36])
    assert (nblocks, ndev, nq, c0, c1, extra) == (1, 2, 3, 64, 64, 0)
    per_dev = raw.reshape(2, 3, 64)
    np.savez_compressed(HERE / "golden_semiglobal_knc.npz", scores=np.concatenate([per_dev[0], per_dev[1]], axis=1),
                        raw=raw, info=np.frombuffer(info, dtype=np.uint8))
    # the same golden as files, plus the text the reference's own convert made of it (tests/test_convert_tool.py)
    shutil.copy(REF / "banded/BGSA_KNC/data/result.txt", HERE / "knc_result.bin")
    shutil.copy(REF / "banded/BGSA_KNC/data/result.txt.info", HERE / "knc_result.bin.info")
    shutil.copy(REF / "banded/BGSA_KNC/data/convert_result.txt", HERE / "knc_convert_result.txt")

    out = {}
    q, s = R.sample_data()
    out["sample_myers_cpu"] = R.reflib("myers_cpu").run(q, s)
    out["sample_myers_sse"] = R.reflib("myers_sse").run(q, s)
    out["sample_bitpal_avx512"] = R.reflib("bitpal_avx512").run(q, s)
    out["sample_bitpal_avx2"] = R.reflib("bitpal_avx2").run(q, s)
    out["sample_banded_k31"] = R.reflib("banded_cpu").run(q, s, e=31)
    # small slices of the BASELINE configs
    for name, n in (("C2", 2048), ("C3", 4096), ("C4", 256), ("C5", 48)):
        qq, ss = synth.make(name, n)
        out[f"{name}_query"] = qq
        out[f"{name}_subjects"] = ss
        if name == "C2":
            out["C2_ref"] = R.reflib("bitpal_avx512").run(qq, ss)
            assert (out["C2_ref"] == R.reflib("bitpal_avx2").run(qq, ss)).all()
            out["C2_myers_ref"] = R.reflib("myers_cpu").run(qq, ss)
        elif name == "C3":
            out["C3_ref"] = R.reflib("banded_cpu").run(qq, ss, e=5)
        elif name == "C4":
            # generator output restated by us inside the reference pipeline (oracle/semiglobal_align_core.c)
            out["C4_ref_restated"] = R.reflib("semiglobal_cpu").run(qq, ss)
            out["C4_myers_ref"] = R.reflib("myers_cpu").run(qq, ss)
        else:
            out["C5_ref"] = R.reflib("bitpal_avx512").run(qq, ss)
    # ragged / N-containing cases through the reference
    rng = np.random.default_rng(99)
    for i, (ql, sl, n) in enumerate([(37, 150, 65), (150, 37, 40), (1, 1, 5), (63, 64, 33), (500, 480, 20)]):
        qq = R.random_rows(rng, 2, ql, with_n=0.03)
        ss = R.random_rows(rng, n, sl, with_n=0.03)
        out[f"rag{i}_query"], out[f"rag{i}_subjects"] = qq, ss
        out[f"rag{i}_myers"] = R.reflib("myers_cpu").run(qq, ss)
        out[f"rag{i}_bitpal"] = R.reflib("bitpal_avx512").run(qq, ss)
    np.savez_compressed(HERE / "ref_vectors.npz", **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
