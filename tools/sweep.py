"""BASELINE.json configs[4]: scaling sweep over batch {64, 256, 1024} x {sync, async} at the GPU count of the launch.

    python tools/sweep.py                      # 1 GPU
    torchrun --nproc-per-node N tools/sweep.py  # not needed: this script launches bench.py itself for every N it is given
    python tools/sweep.py --gpus 1 2 4 8

Each cell is one `bench.py` run (its JSON line is kept whole in gpurun_out/sweep.jsonl); the table printed at the end
has samples/s, microseconds per SGD step and the fraction of the HBM roofline of the dominant kernel."""
import argparse, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, nargs="+", default=[1])
ap.add_argument("--batches", type=int, nargs="+", default=[64, 256, 1024])
ap.add_argument("--modes", nargs="+", default=["sync", "async"])
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
rows = []
port = 29600
for n in args.gpus:
    for mode in args.modes:
        for b in args.batches:
            cmd = [sys.executable]
            if n > 1:
                port += 1
                cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port)]
            cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--mode", mode, "--batch", str(b), "--steps", str(args.steps),
                    "--warmup", str(args.warmup), "--cpu-seconds", "2"]
            if mode == "sync":
                cmd += ["--sgd-steps", str(max(200, 560000 // (b * 4)))]
            r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
            line = next((l for l in r.stdout.splitlines() if l.startswith("{")), None)
            if not line:
                print(f"N={n} {mode} batch {b}: FAILED\n{r.stderr[-500:]}", flush=True)
                continue
            j = json.loads(line)
            with open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "a") as f:
                f.write(line + "\n")
            per = j["config"].get("sgd_steps_per_bench_step") or j["config"].get("updates_per_gpu_per_step")
            rows.append((n, mode, b, j["value"], j["e2e"]["value"], j["ms_per_step"] * 1e3 / per, j["roofline"]["frac"]))
            print(f"N={n} {mode:5s} batch {b:5d}: {j['value']:.4g} samples/s  e2e {j['e2e']['value']:.4g}  "
                  f"{rows[-1][5]:.2f} us/step  roofline {rows[-1][6]:.4f}", flush=True)
print("\n| GPUs | mode | batch | samples/s | e2e samples/s | us per step | HBM roofline frac |\n|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.4g} | {r[4]:.4g} | {r[5]:.2f} | {r[6]:.4f} |")
