#!/usr/bin/env python3
"""Writes profiles/work_model.json: FP64 operations EXECUTED per unit by the shipped Jacobian kernels, from the ncu counters
(smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on) of tracked `ncu --set full` raw-page exports, plus the DRAM bytes
per unit of the same captures.  bench.py reads the file for its roofline object ("instrumented count").
    python profiles/make_work_model.py
Re-run after a kernel change together with a fresh capture (profiles/README.md has the ncu command)."""
import csv
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# family -> (raw csv, units per launch, {label: kernel-name regex}, multiplier per label)
SOURCES = {
    # python profiles/run_kernel.py jvp 16384 2, launches 7-9: one full chunk of 2^20 units
    "chain6": ("r02/r02_c2_jvp_raw.csv", 1048576, {"step_stages": "k_step_stages", "stage_derivs": "k_stage_derivs", "chain_rule": "k_chain_rule"}, 1),
    # python profiles/run_kernel.py jvp 4096 1 pilz6x2c 100, launches 9-16: coupled-fatigue pre kernel, both chains, post kernel
    "forest12x6": ("r02/r02_c3_jvp_raw.csv", 409600, {"step_stages": "k_step_stages", "stage_derivs": "k_stage_derivs", "chain_rule": "k_chain_rule",
                                                     "couple": "k_couple"}, 1),
    # MPCF_TREE_CHAIN=scalar python profiles/run_kernel.py jvp 1024 1 humanoid37 40, launches 9-12: the tree pipeline with the DFMA
    # chain kernel on its first chunk.  The default (tensor-core) chain kernel runs the same recursion on operands padded to
    # 40 x 80 x 64-column slabs; its padding is not work, so the algorithmic count of C4 is the unpadded DFMA variant's
    # (units of the chunk: grid of k_tree_stages x 128 threads).  The tensor-core capture itself: r02/r02_c4_tc_raw.csv.
    "generic64": ("r02/r02_c4_tree_raw.csv", None, {"tree_stages": "k_tree_stages", "tree_derivs": "k_tree_derivs", "tree_factor": "k_tree_factor", "tree_chain": "k_tree_chain"}, 1),
}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    out = {}
    for fam, (path, units, labels, mult) in SOURCES.items():
        full = os.path.join(HERE, path)
        if not os.path.exists(full):
            continue
        rows = list(csv.reader(open(full)))
        hdr, data = rows[0], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        kernels, dram = {}, 0.0
        if units is None:  # thread-per-unit first kernel: units = its grid x block (exact up to the last block's padding)
            r0 = next(r for r in data if re.search(next(iter(labels.values())), r[ix["Kernel Name"]]))
            units = int(num(r0[ix["launch__grid_size"]]) * num(r0[ix["launch__block_size"]]))
        for r in data:
            name = r[ix["Kernel Name"]]
            label = next((lb for lb, rx in labels.items() if re.search(rx, name)), None)
            if label is None:
                continue
            cyc = num(r[ix["smsp__cycles_elapsed.avg"]])
            k = kernels.setdefault(label, {"dadd": 0.0, "dmul": 0.0, "dfma": 0.0, "launches": 0, "kernel": name.split("(")[0]})
            for op in ("dadd", "dmul", "dfma"):
                k[op] += num(r[ix["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op]]) * cyc / units
            k["launches"] += 1
            dram += (num(r[ix["dram__bytes_read.sum"]]) + num(r[ix["dram__bytes_write.sum"]])) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(
                rows[1][ix["dram__bytes_read.sum"]], 1) / units
        # a capture that holds the same kernel several times for the SAME units (repeated passes) is averaged; the two chains of
        # a forest are different work on the same units and are summed
        reps = 1  # every capture holds exactly one pass over its units (the two chains of a forest are summed)
        for k in kernels.values():
            for op in ("dadd", "dmul", "dfma"):
                k[op] = round(k[op] / reps * mult)
        out[fam] = {"kernels": kernels, "dram_bytes_per_unit": round(dram / reps), "source": "profiles/" + path, "units_per_launch": units}
    # the default chain kernel of the tree pipeline runs on the FP64 tensor cores: DMMA.8x8x4 warp instructions per unit from
    # the tensor sub-pipe's active cycles (16 per instruction and SM sub-partition), for the tensor-pipe utilisation in bench.py
    tc = os.path.join(HERE, "r02/r02_c4_tc_raw.csv")
    if os.path.exists(tc) and "generic64" in out:
        rows = list(csv.reader(open(tc)))
        hdr, data = rows[0], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        r0 = next(r for r in data if "k_tree_stages" in r[ix["Kernel Name"]])
        units = int(num(r0[ix["launch__grid_size"]]) * num(r0[ix["launch__block_size"]]))
        dcol = next(h for h in hdr if h.endswith("smsp__pipe_tensor_subpipe_dmma_cycles_active.avg"))
        info = {"source": "profiles/r02/r02_c4_tc_raw.csv", "units_per_launch": units, "kernels": {}}
        for r in data:
            name = r[ix["Kernel Name"]].split("(")[0]
            smsp = 4 * num(r[ix["device__attribute_multiprocessor_count"]])
            dmma = num(r[ix[dcol]]) * smsp / 16.0 / units
            cyc = num(r[ix["smsp__cycles_elapsed.avg"]])
            k = {"dmma_warp_instr": round(dmma), "ms": num(r[ix["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(
                rows[1][ix["gpu__time_duration.sum"]], 1.0)}
            for op in ("dadd", "dmul", "dfma"):
                k[op] = round(num(r[ix["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op]]) * cyc / units)
            info["kernels"][name] = k
        out["generic64"]["tensor_core_pipeline"] = info
    with open(os.path.join(HERE, "work_model.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    for fam, v in out.items():
        flop = sum(k["dadd"] + k["dmul"] + 2 * k["dfma"] for k in v["kernels"].values())
        inst = sum(k["dadd"] + k["dmul"] + k["dfma"] for k in v["kernels"].values())
        print("%-12s %9d FLOP/unit  %9d FP64 instr/unit  %7d DRAM B/unit   (%s)" % (fam, flop, inst, v["dram_bytes_per_unit"], v["source"]))


if __name__ == "__main__":
    main()
