#!/bin/bash
# Probe the GPU box for the real reference dependencies (PyBullet + gym<0.22).
# prove whether the physics can be pinned against PyBullet.
# Output goes to gpurun_out/probe_reference_deps.log (copied to profiles/ by hand).
out=gpurun_out/probe_reference_deps.log
mkdir -p gpurun_out
{
  echo "== date: $(date -u +%FT%TZ)  host: $(hostname)  cores: $(nproc)"
  echo "== python: $(which python) $(python --version 2>&1)"
  for m in pybullet gym gymnasium sacred mujoco dm_control pybullet_envs pybullet_data; do
    python - <<PY 2>&1
try:
    import $m
    print("import $m: OK", getattr($m, "__version__", "?"), getattr($m, "__file__", "?"))
except Exception as e:
    print("import $m: FAIL", type(e).__name__, e)
PY
  done
  echo "== pip download / install attempts (expect: no network)"
  timeout 60 python -m pip download --no-deps -d /tmp/pb pybullet 2>&1 | tail -3
  timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse pybullet 2>&1 | tail -3
  timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse 'gym<0.22' 2>&1 | tail -3
  echo "== wheelhouse:"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "bullet|gym|mujoco" || echo "(no bullet/gym/mujoco wheels in /opt/wheelhouse)"
  echo "== filesystem search for Bullet:"
  find / -xdev \( -iname "*pybullet*" -o -iname "btMultiBody*" -o -iname "libBullet*" \) -not -path "*/proc/*" 2>/dev/null | grep -v "$GRAFT_REPO_ROOT" | head -20 || true
  echo "(end of search)"
  echo "== baseline/_ref:"; ls -la baseline/_ref 2>&1 | head
  echo "== /root/reference:"; ls /root/reference 2>&1 | head -3
  echo "== network:"; timeout 10 python - <<'PY' 2>&1
import socket
try:
    socket.create_connection(("pypi.org", 443), timeout=5); print("pypi.org reachable")
except Exception as e:
    print("pypi.org unreachable:", type(e).__name__, e)
PY
} > "$out" 2>&1
cat "$out"
