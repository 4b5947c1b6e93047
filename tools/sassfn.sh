#!/bin/bash
# usage: tools/sassfn.sh <lib.so> <kernel-name-substring>  -> the kernel's SASS, one instruction per line (address, text)
cuobjdump -sass "$1" 2>/dev/null | awk -v k="$2" '/Function : /{f=(index($0,k)>0)} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
