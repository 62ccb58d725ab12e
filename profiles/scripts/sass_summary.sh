#!/bin/bash
# Per-kernel SASS evidence of libgr_b200.so (runs on the CPU build box: cuobjdump only needs the .so)
SO=gnn-recommendations_b200/libgr_b200.so
OUT=profiles/sass_summary.txt
{
echo "# cuobjdump -sass $SO  (built by make -C gnn-recommendations_b200/csrc; nvcc $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# per kernel: instruction count and counts of the mnemonics that prove the Blackwell paths"
echo "# UTCHMMA = tcgen05.mma (tf32), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, LDGSTS = cp.async, SYNCS = mbarrier,"
echo "# UTMALDG/UTMASTG/UBLKCP = TMA (cp.async.bulk*), ST.E.*SYS / LD.E.*SYS = system-scope (peer) accesses, DFMA = fp64 (matrix exp)"
cuobjdump -sass $SO | awk '
/Function :/ { if (name != "") flush(); name=$3; n=0; delete c; next }
/^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//  { n++; for (k in pat) if ($0 ~ pat[k]) c[k]++ }
function flush() { printf "%-90s instr=%-6d", name, n; for (k in pat) if (c[k] > 0) printf " %s=%d", k, c[k]; printf "\n" }
BEGIN { pat["UTCHMMA"]="UTCHMMA"; pat["LDTM"]="LDTM"; pat["UTCBAR"]="UTCBAR"; pat["LDGSTS"]="LDGSTS"; pat["SYNCS"]="SYNCS"; pat["UTMALDG"]="UTMALDG"; pat["UTMASTG"]="UTMASTG"; pat["UBLKCP"]="UBLKCP"; pat["RED"]="RED\\."; pat["DFMA"]="DFMA"; pat["FFMA"]="FFMA"; pat["SHFL"]="SHFL"; pat["MUFU_EX2"]="MUFU.EX2"; pat["SYS"]="\\.SYS" }
END { flush() }' | c++filt | sed 's/gr:://g' | sort
} > $OUT
wc -l $OUT
