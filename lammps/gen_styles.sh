#!/bin/sh
# Regenerates LAMMPS' style_*.h registries for the files present in the current directory
# (same grep rule as the reference's src/Make.sh:12-75, restated).
gen () {  # macro prefix name
  out=style_$3.h
  : > $out
  for f in `grep -sl $1 $2*.h | sort`; do
    case $f in style_*) ;; *) echo "#include \"$f\"" >> $out;; esac
  done
}
gen ANGLE_CLASS angle_ angle
gen ATOM_CLASS atom_vec_ atom
gen BODY_CLASS body_ body
gen BOND_CLASS bond_ bond
gen COMMAND_CLASS "" command
gen COMPUTE_CLASS compute_ compute
gen DIHEDRAL_CLASS dihedral_ dihedral
gen DUMP_CLASS dump_ dump
gen FIX_CLASS fix_ fix
gen IMPROPER_CLASS improper_ improper
gen INTEGRATE_CLASS "" integrate
gen KSPACE_CLASS "" kspace
gen MINIMIZE_CLASS min_ minimize
gen PAIR_CLASS pair_ pair
gen READER_CLASS reader_ reader
gen REGION_CLASS region_ region
