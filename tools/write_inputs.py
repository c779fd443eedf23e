#!/usr/bin/env python
"""Write bench.py's config-2 workload (same seed, same arrays) as little-endian binary files, so that a box that has Julia can
time the REAL reference on exactly the inputs the GPU arm used (same binary inputs for oracle, GPU and Julia).

    python tools/write_inputs.py OUTDIR [nsteps]     ->  OUTDIR/{I,J,V}.bin (Int64, Int64, Float64), OUTDIR/batch_%03d_{i,j,v}.bin,
                                                         OUTDIR/x.bin, OUTDIR/meta.txt
    julia julia/bench_reference.jl OUTDIR            ->  one JSON line in bench.py's `--impl reference` format
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    out = sys.argv[1]
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    os.makedirs(out, exist_ok=True)
    (I, J, V), batches, x = bench.make_workload(nsteps)
    I.astype("<i8").tofile(os.path.join(out, "I.bin"))
    J.astype("<i8").tofile(os.path.join(out, "J.bin"))
    V.astype("<f8").tofile(os.path.join(out, "V.bin"))
    x.astype("<f8").tofile(os.path.join(out, "x.bin"))
    for s, (bi, bj, bv) in enumerate(batches):
        bi.astype("<i8").tofile(os.path.join(out, f"batch_{s:03d}_i.bin"))
        bj.astype("<i8").tofile(os.path.join(out, f"batch_{s:03d}_j.bin"))
        bv.astype("<f8").tofile(os.path.join(out, f"batch_{s:03d}_v.bin"))
    with open(os.path.join(out, "meta.txt"), "w") as f:
        f.write(f"{bench.M_ROWS} {bench.N_COLS} {len(I)} {bench.BATCH} {nsteps}\n")
    print(f"wrote {len(I)} initial entries, {nsteps} batches of {bench.BATCH} to {out}")


if __name__ == "__main__":
    main()
