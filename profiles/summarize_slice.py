"""Summarise the ncu launch-list slice of profiles/launch_slice.py: encoder micro-batch x 8 + decode steps x 30 = step."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    order = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"])
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        order.append((name, v))
    # the decode slice starts with LayerNorm -> QKV GEMM -> self-attention (append) of layer 0
    first_attn = next(i for i, (n, _) in enumerate(order) if "attention_decode" in n)
    dec_start = first_attn - (1 if "decode_chain" in order[first_attn - 1][0] else 2)  # fused: one chain launch before it
    enc, dec = order[:dec_start], order[dec_start:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in enc:
        agg[n][0] += 8
        agg[n][1] += 8 * v
    for n, v in dec:
        agg[n][0] += 30
        agg[n][1] += 30 * v
    tot = sum(v[1] for v in agg.values())
    print(f"step estimate (cold-cache, serialised): {tot / 1000:.1f} ms over {sum(v[0] for v in agg.values())} kernels; "
          f"encoder slice {sum(v for _, v in enc) / 1000:.2f} ms x 8, decode slice {sum(v for _, v in dec) / 1000:.3f} ms x 30")
    print("| kernel | total ms | share | launches | avg us |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {t / 1000:.3f} | {100 * t / tot:.1f}% | {n} | {t / n:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
