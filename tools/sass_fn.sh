#!/bin/bash
# usage: tools/sass_fn.sh <object> <mangled-name-substring>  -> SASS of the first matching function on stdout
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function : /{p = index($0, pat) > 0 ? 1 : 0} p{print}'
