#!/bin/bash
# usage: tools/spill_sites.sh OBJECT MANGLED_KERNEL_PREFIX  -- local-memory (spill) instructions of one kernel by source line
obj=$1; fun=$2; tmp=$(mktemp -d)
( cd $tmp && cuobjdump -xelf all $OLDPWD/$obj >/dev/null 2>&1 && nvdisasm -g *.cubin > all.dis )
awk -v f="$fun" '/^\s*\.text\./{on = index($0, f) > 0} on' $tmp/all.dis > $tmp/k.dis
awk '/\/\/## File/{f=$0; sub(/.*csrc\//,"",f); sub(/.*include\//,"",f)} /STL|LDL/{print f}' $tmp/k.dis | sort | uniq -c | sort -k1nr | head -${3:-30}
rm -rf $tmp
