#!/bin/bash
# Records whether lalsimulation is reachable on the GPU box (SURVEY §7 first action).
out=gpurun_out/lal_probe.txt
mkdir -p gpurun_out
{
echo "== date: $(date -u)"; echo "== host: $(hostname)"; nproc; free -g | head -2
echo "== import lalsimulation"; python -c "import lalsimulation; print('lalsimulation', lalsimulation.__version__)" 2>&1 | tail -1
echo "== import lal"; python -c "import lal; print('lal', lal.__version__)" 2>&1 | tail -1
echo "== import astropy / h5py / dynesty"; for m in astropy h5py dynesty pycbc; do python -c "import $m; print('$m', $m.__version__)" 2>&1 | tail -1; done
echo "== pip download lalsuite"; timeout 40 python -m pip download lalsuite --no-deps -d /tmp/lal_dl 2>&1 | tail -3
echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i "lal\|astropy\|h5py" || echo "no lal/astropy/h5py wheel in /opt/wheelhouse"
echo "== baseline/_ref"; ls baseline/_ref 2>&1 | head
echo "== files named *lalsim* outside the repo"; find / -xdev -iname "*lalsim*" -not -path "/proc/*" -not -path "$GRAFT_REPO_ROOT/*" 2>/dev/null | head -5; echo "(end)"
echo "== conda"; which conda mamba 2>&1 | tail -2
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
} > $out 2>&1
cat $out
