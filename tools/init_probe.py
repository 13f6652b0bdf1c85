import subprocess, os, time, json, sys
ROOT="/root/repo"
exe=os.path.join(ROOT,"build","dropin")
for tag, extra in (("all_visible", {}), ("visible_0", {"CUDA_VISIBLE_DEVICES":"0"}), ("all_visible_again", {}), ("visible_0_again", {"CUDA_VISIBLE_DEVICES":"0"}), ("lazy_off", {"CUDA_MODULE_LOADING":"EAGER"})):
    env=dict(os.environ, RT_B200_TIMING="1", **extra)
    t0=time.time()
    r=subprocess.run([exe,"quads","/dev/shm/x.ppm","64","4"],env=env,capture_output=True,text=True)
    wall=time.time()-t0
    parts=[json.loads(l[14:]) for l in r.stderr.splitlines() if l.startswith("RTB200_TIMING ")]
    print(tag, round(wall*1e3,1), parts[-1]["init_ms"] if parts else r.stderr[-200:])
print(subprocess.run(["nvidia-smi","-L"],capture_output=True,text=True).stdout)
print(subprocess.run(["nvidia-smi","--query-gpu=persistence_mode","--format=csv"],capture_output=True,text=True).stdout)
