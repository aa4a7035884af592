import os, subprocess, sys
ROOT="/root/repo"
for name in sys.argv[1:]:
    env=dict(os.environ)
    if name!="base": env["VQB_LIB_PATH"]=f"{ROOT}/vq-vae-transformer-arc-welding_b200/ab_{name}.so"
    r=subprocess.run([sys.executable, f"{ROOT}/tools/bwd_time.py"],env=env,capture_output=True,text=True)
    print(name, [l for l in r.stdout.splitlines() if l.startswith("bwd")])
