"""Aggregates the [trace] lines DMRGX_TRACE=1 writes to stderr: total ms and count per mark."""
import collections
import re
import sys

agg = collections.defaultdict(lambda: [0, 0.0])
for line in open(sys.argv[1]):
    m = re.match(r"\[trace\] (\S+) ([0-9.]+) ms", line)
    if m:
        agg[m.group(1)][0] += 1
        agg[m.group(1)][1] += float(m.group(2))
    elif line.startswith("[trace] allocator"):
        print(line.strip())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-32s n=%6d  total %10.1f ms  mean %8.3f ms" % (k, c, t, t / c))
