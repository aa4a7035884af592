"""GPU box: tools/bwd_time.py under experiment builds of the library (see ab_build.py)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in sys.argv[1:]:
    env=dict(os.environ)
    if name!="base": env["VQB_LIB_PATH"]=f"{ROOT}/vq-vae-transformer-arc-welding_b200/ab_{name}.so"
    r=subprocess.run([sys.executable, f"{ROOT}/tools/bwd_time.py"],env=env,capture_output=True,text=True)
    print(name, [l for l in r.stdout.splitlines() if l.startswith("bwd")])
