#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the small JSON kept under profiles/ (run here, no GPU needed):

    python tests/ncu_summary.py gpurun_out/prof_nt.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_tc_kernels_summary.json

One entry per profiled launch: duration, DRAM bytes read / written, L2 (lts) bytes and throughput, tensor-pipe activity,
SM clock, registers.  `bench.py` reads `prof_nt` / `prof_tn` from the newest `*ncu_tc_kernels_summary.json` for
`roofline.traffic`.  Not a pytest file.
"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = {
    "duration_us": ("gpu__time_duration.sum", None),                         # -> us
    "dram_read_MB": ("dram__bytes_read.sum", None),
    "dram_write_MB": ("dram__bytes_write.sum", None),
    "lts_read_MB": ("lts__t_sectors_op_read.sum", 32e-6),                    # sectors -> MB
    "lts_write_MB": ("lts__t_sectors_op_write.sum", 32e-6),
    "l2_throughput_pct": ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l2_hit_pct": ("lts__t_sector_hit_rate.pct", 1.0),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "tensor_pipe_active_pct": ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l2_to_sm_read_TBps": ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", None),          # crossbar -> L1/shared: the operand stream
    "l2_to_sm_read_GB": ("l1tex__m_xbar2l1tex_read_bytes.sum", None),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "issue_active_pct": ("sm__inst_issued.avg.pct_of_peak_sustained_active", 1.0),
    "registers": ("launch__registers_per_thread", 1.0),
    "sm_clock_ghz": ("sm__cycles_elapsed.avg.per_second", None),
}
UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0,
              "cycle/nsecond": 1.0, "cycle/second": 1e-9, "cycle/usecond": 1e-3,
              "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
              "byte/s": 1e-12, "Kbyte/s": 1e-9, "Mbyte/s": 1e-6, "Gbyte/s": 1e-3, "Tbyte/s": 1.0}
GB_KEYS = {"l2_to_sm_read_GB": {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    res = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        e = {"kernel": r[names.index("Kernel Name")], "grid": r[names.index("Grid Size")]}
        for key, (metric, scale) in WANT.items():
            idx = [i for i, n in enumerate(names) if n == metric or n.endswith("." + metric)]
            if not idx:
                continue
            try:
                v = float(r[idx[0]].replace(",", ""))
            except ValueError:
                continue
            table = GB_KEYS.get(key, UNIT_SCALE)
            e[key] = v * (scale if scale is not None else table.get(units[idx[0]], 1.0))
        res.append(e)
    return res


def main():
    out = {}
    for rep in sys.argv[1:]:
        key = "prof_tn" if "tn" in os.path.basename(rep) else ("prof_nt" if "nt" in os.path.basename(rep) else os.path.basename(rep).split(".")[0])
        out.setdefault(key, []).extend(load(rep))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
