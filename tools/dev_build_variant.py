"""Dev tool: build libes_b200_<name>.so with an alternative rollout_tc.cu (A/B timing of kernel experiments in one GPU session).
usage: python tools/dev_build_variant.py <name> <path/to/rollout_tc_variant.cu> [-DFLAG ...]"""
import os, shutil, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from es_pytorch_b200 import build as b
name, src = sys.argv[1], sys.argv[2]
with tempfile.TemporaryDirectory() as d:
    shutil.copytree(b.CSRC, os.path.join(d, 'es_pytorch_b200', 'csrc'))
    shutil.copytree(os.path.join(os.path.dirname(b.HERE), 'include'), os.path.join(d, 'include'))
    shutil.copy(src, os.path.join(d, 'es_pytorch_b200', 'csrc', 'rollout_tc.cu'))
    out = os.path.join(b.HERE, f'libes_b200_{name}.so')
    cmd = [b.nvcc_path()] + b.NVCC_FLAGS + sys.argv[3:] + ['-o', out] + [os.path.join(d, 'es_pytorch_b200', 'csrc', s) for s in b._sources()]
    subprocess.run(cmd, check=True)
print(out)
