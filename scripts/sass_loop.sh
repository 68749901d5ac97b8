#!/bin/bash
# usage: scripts/sass_loop.sh <object> <mangled kernel name substring>  -> instruction mix of the whole kernel + regs
OBJ=$1; PAT=$2
FN=$(cuobjdump -sass $OBJ | grep "Function :" | grep "$PAT" | head -1 | awk '{print $3}')
echo "kernel: $FN"
cuobjdump -sass -fun "$FN" $OBJ | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/t/k.lst
wc -l /tmp/t/k.lst
