#!/bin/sh
# usage: tools/sass_of.sh <object> <substring of the function name>  -> SASS without the encoding lines
obj="$1"; pat="$2"
fn=$(cuobjdump -sass "$obj" | grep "Function :" | grep "$pat" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$fn" "$obj" | grep -v '^\s*/\* 0x'
