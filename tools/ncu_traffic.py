"""DRAM traffic of the dominant kernels per launch, from an ncu metrics capture of bench.py, written to
profiles/ncu_traffic.json (bench.py copies it into roofline.traffic; nothing there is typed in by hand).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:'gemm_tc_kernel|gram_|conv3x3_halo|stem_conv' -c 1200 --csv --log-file gpurun_out/ncu_traffic.csv \
        python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-graph --no-pipeline --no-roofline
    python tools/ncu_traffic.py gpurun_out/ncu_traffic.csv 5 > profiles/ncu_traffic.json      (5 = steps in the capture)"""
import collections
import csv
import json
import re
import sys


def main(path, steps):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui, ii = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[ii], {"name": re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")})
        v = float(r[vi].replace(",", ""))
        u = r[ui].lower()
        if "byte" in u:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        elif u in ("ns", "nsecond"):
            v /= 1e3
        elif u in ("ms", "msecond"):
            v *= 1e3
        d[r[mi]] = v
    launches = list(per.values())
    n = len(launches) // steps
    last = launches[-n:]                                   # the last step of the capture
    fam = [x for x in last if x["name"].startswith(("gemm_tc_kernel", "gram_stats", "conv3x3_halo"))]
    post = [x for x in fam if re.search(r"gemm_tc_kernel<256, *\(?EpiMode\)?3|gemm_tc_kernel<256, 3", x["name"])]

    def tot(xs):
        return sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in xs)
    out = {"source": path, "steps_in_capture": steps, "launches_per_step": n,
           "family": {"kernels": "gemm_tc_kernel + gram_stats_kernel + conv3x3_halo_kernel", "launches": len(fam),
                      "dram_bytes_per_step": tot(fam), "dram_bytes_per_launch": tot(fam) / max(len(fam), 1),
                      "time_us_per_step": sum(x.get("gpu__time_duration.sum", 0) for x in fam),
                      "note": "dram__bytes_read.sum + dram__bytes_write.sum over every tcgen05 launch of one eager step, "
                              "ncu metrics pass (cold-cache, serialised), averaged per launch"},
           "hbm_member": {"kernels": "gemm_tc_kernel<256, EPI_POST, ...> (conv3 + BN3 + shortcut + ReLU)", "launches": len(post),
                          "dram_bytes_per_step": tot(post), "dram_bytes_per_launch": tot(post) / max(len(post), 1),
                          "time_us_per_step": sum(x.get("gpu__time_duration.sum", 0) for x in post),
                          "note": "same capture, the EPI_POST launches only"}}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
