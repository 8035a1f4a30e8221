"""profiles/ncu_traffic.json from an `ncu --set full` raw CSV (ncu -i x.ncu-rep --page raw --csv > profiles/<name>.csv):
per kernel-name regex, mean dram__bytes_read.sum + dram__bytes_write.sum per launch, with the sha256 of the CSV so
that bench.py can tell whether the number it reports still belongs to the committed capture.

    python tools/ncu_traffic.py profiles/r2_ncu_prop_raw.csv spmm_layer_kernel/ml-20m spmm_layer_kernel [more key regex ...]
"""
import csv, hashlib, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    src = sys.argv[1]
    pairs = list(zip(sys.argv[2::2], sys.argv[3::2]))
    rows = list(csv.reader(open(src)))
    h, units = rows[0], rows[1]
    ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
    for key, pat in pairs:
        vals = [float(r[ri].replace(",", "")) * UNIT[units[ri]] + float(r[wi].replace(",", "")) * UNIT[units[wi]]
                for r in rows[2:] if re.search(pat, r[ki])]
        if not vals:
            print("no launch matches", pat)
            continue
        table[key] = {"dram_bytes_per_launch": int(sum(vals) / len(vals)), "launches": len(vals), "kernel": pat,
                      "source": os.path.relpath(src, os.path.join(ROOT, "profiles")), "sha256": sha,
                      "captured_with": "ncu --set full --clock-control none"}
        print(key, table[key])
    json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
