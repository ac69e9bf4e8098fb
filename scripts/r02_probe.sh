#!/bin/bash
# Round-2 first GPU call: Fortran compiler probe on the GPU box, baseline GPU tests, fresh ncu capture of kick_flat_kernel.
mkdir -p gpurun_out
{
  echo "## which"; for c in gfortran gfortran-13 gfortran-12 f95 f77 ifx ifort nvfortran pgfortran flang flang-new lfortran caf mpif90 mpifort h5fc nf-config; do printf "%s: " $c; command -v $c || echo absent; done
  echo "## find f951 / compiler binaries"; find / \( -name f951 -o -name 'gfortran*' -o -name 'flang*' -o -name nvfortran -o -name ifx -o -name 'lfortran*' \) -not -path '/proc/*' 2>/dev/null | head -20
  echo "## conda"; (conda list 2>&1 | grep -i fortran) || echo "no conda / no fortran package"
  echo "## dpkg"; (dpkg -l 2>/dev/null | grep -i -E 'fortran|flang' ) || echo "no dpkg fortran package"
  echo "## cpu"; nproc; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Core|Socket'
} > gpurun_out/r02_fortran_probe.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest0.log 2>&1; echo "pytest rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kick_flat_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_kick_flat_base python scripts/kick_bench.py 100000 1 > gpurun_out/r02_ncu0.log 2>&1; echo "ncu rc=$?"
python scripts/kick_bench.py 100000 5 > gpurun_out/r02_kickbench0.log 2>&1
tail -3 gpurun_out/r02_pytest0.log; cat gpurun_out/r02_kickbench0.log; cat gpurun_out/r02_fortran_probe.txt
