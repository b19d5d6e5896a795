#!/bin/bash
# usage: tools/sass_fn.sh <lib.so> <mangled-kernel-name>  -> clean SASS listing of one kernel
cuobjdump -sass "$1" | awk -v fn="$2" '/Function : /{f=($3==fn)} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/ +;/;/'
