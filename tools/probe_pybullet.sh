#!/bin/bash
# Probe the GPU box for pybullet: record what is there, verbatim.
O=gpurun_out/r02_pybullet_probe.txt
{
echo "== date: $(date -u +%FT%TZ) host: $(hostname)"
echo "== python -c 'import pybullet'"; python -c 'import pybullet; print(pybullet.__file__); print(pybullet.getAPIVersion())' 2>&1
echo "== python -c 'import pybullet_data'"; python -c 'import pybullet_data; print(pybullet_data.getDataPath())' 2>&1
echo "== pip list | grep -i bullet"; python -m pip list 2>/dev/null | grep -i bullet || echo "(none)"
echo "== find / -iname '*bullet*' (pruned /proc /sys)"; find / \( -path /proc -o -path /sys \) -prune -o -iname '*bullet*' -print 2>/dev/null | head -20 || true
echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'bullet|fcl|trimesh|coal|hpp' || echo "(no bullet/fcl/trimesh wheel)"
echo "== other collision libraries"; for m in fcl trimesh coal hppfcl mujoco pinocchio; do python -c "import $m; print('$m', $m.__file__)" 2>&1 | tail -1; done
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
echo "== cpu"; nproc; grep -m1 'model name' /proc/cpuinfo
} > $O 2>&1
cat $O
