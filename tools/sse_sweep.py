import os, sys, subprocess, json
for w in (1, 2, 4, 8, 16):
    env = dict(os.environ, GOOFY_B200_SSE_WAVES=str(w))
    out = subprocess.run([sys.executable, "tools/bench_next_rows.py", "--steps", "40"], capture_output=True, text=True, env=env).stdout
    d = json.loads(out)["results"]
    print("waves", w, "sse dxt1 %.0f etc1 %.0f GB/s" % (d["block_sse_dxt1"]["gb_per_s"], d["block_sse_etc1"]["gb_per_s"]), flush=True)
