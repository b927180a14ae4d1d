"""Two ranks of bench.py on two GPUs, rank 0 under compute-sanitizer memcheck (torchrun cannot wrap a single rank)."""
import os
import subprocess
import sys

args = ["bench.py", "--quick", "--gpus", "2", "--steps", "6", "--warmup", "5"]
procs = []
for rank in range(2):
    env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29577",
               OMP_NUM_THREADS="1")
    cmd = [sys.executable] + args
    if rank == 0:
        cmd = ["compute-sanitizer", "--tool", "memcheck", "--print-limit", "8", "--error-exitcode", "7"] + cmd
    out = open("gpurun_out/san_rank%d.log" % rank, "w")
    procs.append(subprocess.Popen(cmd, env=env, stdout=out, stderr=subprocess.STDOUT))
for p in procs:
    try:
        p.wait(timeout=int(sys.argv[1]) if len(sys.argv) > 1 else 500)
    except subprocess.TimeoutExpired:
        p.kill()
print("exit codes", [p.returncode for p in procs])
