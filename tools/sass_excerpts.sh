#!/bin/bash
# SASS evidence for profiles/: the TMA instructions of the field solver, the opcode mix of the hot push kernels
# (static, from the built objects; the dynamic mix is in the ncu summaries)
cd "$(dirname "$0")/../epoch_b200/csrc"
out=../../profiles/r02_sass_excerpts.txt
{
echo "# cuobjdump -sass of the objects that make libepoch_b200.so (sm_100a), $(nvcc --version | tail -2 | head -1)"
echo
echo "## fdtd_tma.o: TMA tile loads (UTMALDG) and their mbarrier handshakes (SYNCS) in k_fdtd_tma_2d / _3d"
cuobjdump -sass fdtd_tma.o | grep -E "Function :|UTMALDG|SYNCS|UTMAPF|ELECT" | sed 's/^[[:space:]]*//' | cut -c1-150
echo
for fn in push_slots_2dILi8ELi3ELb0ELb0 push_bag_3dILb0; do
  sym=$(cuobjdump -sass push_fast.o | grep "Function :" | grep "$fn" | head -1 | awk '{print $3}')
  echo "## push_fast.o: $sym -- static opcode counts"
  cuobjdump -sass -fun "$sym" push_fast.o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/;//' | sort | uniq -c | sort -rn | head -40 | awk '{printf "%6d %s\n", $1, $2}'
  echo
  echo "## ... its global-memory instructions (row loads / stores are 64-bit, one per component; inbox entries 128-bit)"
  cuobjdump -sass -fun "$sym" push_fast.o | grep -E "LDG|STG|RED|ATOMG|LDGSTS" | sed -E 's/^\s+//' | awk '{ $1=""; print }' | sed 's/^ //' | sort | uniq -c | sort -rn | head -30
  echo
  echo "## ... its shared-memory atomics (FP64 add in shared memory = compare-and-swap loop)"
  cuobjdump -sass -fun "$sym" push_fast.o | grep -cE "ATOMS.CAST.SPIN" | awk '{print $1 " ATOMS.CAST.SPIN.64 sites"}'
  echo
done
cuobjdump --dump-resource-usage push_fast.o | grep -A1 -E "push_slots_2dILi8ELi3ELb0ELb0|push_bag_3dILb0" | grep -E "Function|REG" | sed 's/^ *//'
} > $out 2>&1
wc -l $out
