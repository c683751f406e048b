"""Turns the CSV of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
--csv --log-file X python bench.py --steps 2 --warmup 3` into the committed launch list (one row per kernel launch)
and refreshes profiles/conv_traffic.json (DRAM bytes per conv launch of the LAST step, read by bench.py's `roofline.traffic`).
Usage: python tools/ncu_launches.py gpurun_out/launches.csv profiles/r01_ncu_launches.csv [profiles/conv_traffic.json]"""
import csv
import json
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ci = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
    launches = {}
    order = []
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        k = int(r[ci["ID"]])
        if k not in launches:
            name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("io::", "")
            launches[k] = dict(kernel=name, grid=r[ci["Grid Size"]], block=r[ci["Block Size"]])
            order.append(k)
        v = float(r[ci["Metric Value"]].replace(",", ""))
        m, u = r[ci["Metric Name"]], r[ci["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        launches[k][m] = v * scale
    with open(dst, "w") as f:
        f.write("idx,kernel,grid,block,duration_us,dram_read_MB,dram_write_MB\n")
        for i, k in enumerate(order):
            L = launches[k]
            f.write('%d,%s,"%s","%s",%.2f,%.2f,%.2f\n' % (i, L["kernel"], L["grid"], L["block"],
                    L.get("gpu__time_duration.sum", 0), L.get("dram__bytes_read.sum", 0), L.get("dram__bytes_write.sum", 0)))
    # a step = the launches from one gather launch up to the next; the capture (-c N) may cut the last one short, so take
    # the last step that is as long as the longest one
    names = [launches[k]["kernel"] for k in order]
    starts = [i for i, n in enumerate(names) if n.startswith("gather_patch")] + [len(order)]
    segs = [(starts[i], starts[i + 1]) for i in range(len(starts) - 1)]
    full = max(b - a for a, b in segs)
    a, b = [sg for sg in segs if sg[1] - sg[0] == full][-1]
    step = [launches[k] for k in order[a:b]]
    conv = [L for L in step if L["kernel"].startswith(("conv_", "stem_"))]   # stem_pool_kernel is conv1 + pool
    tot = sum(L.get("gpu__time_duration.sum", 0) for L in step)
    ct = sum(L.get("gpu__time_duration.sum", 0) for L in conv)
    out = dict(dram_bytes_per_launch=sum(L.get("dram__bytes_read.sum", 0) + L.get("dram__bytes_write.sum", 0) for L in conv) * 1e6 / len(conv),
               conv_launches_per_step=len(conv), conv_share_of_step_ncu=ct / tot, step_kernel_time_us_ncu=tot,
               source="%s (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                      "python bench.py --steps 2 --warmup 3; last step)" % dst)
    print(json.dumps(out, indent=1))
    by = {}
    for L in step:
        a = by.setdefault(L["kernel"], [0, 0.0])
        a[0] += 1; a[1] += L.get("gpu__time_duration.sum", 0)
    for n, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print("%-40s %3d launches %9.1f us %5.1f %%" % (n, c, t, 100 * t / tot))
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
