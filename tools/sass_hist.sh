#!/bin/bash
# usage: tools/sass_hist.sh <object> <function-substring>  -> opcode histogram + spill sites of the matching kernel
OBJ=$1; PAT=$2
FN=$(cuobjdump -sass $OBJ | grep "Function :" | grep "$PAT" | head -1 | awk '{print $3}')
echo "function: $FN"
cuobjdump -sass -fun "$FN" $OBJ > /tmp/sass_fn.txt
grep -E "^\s+/\*[0-9a-f]+\*/" /tmp/sass_fn.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -${3:-28}
echo "spill sites:"; grep -nE "STL|LDL" /tmp/sass_fn.txt | awk '{print $1,$2,$3,$4,$5}' | head -40
