#!/bin/sh
# Prints the .cc files of the reference's w2rap-contigger target (CMakeLists.txt:55,361,398), one per line,
# relative to the reference root. Reads CMakeLists.txt as a list; cmake itself is never run.
REF="${1:-/root/reference}"
awk '
/^add_library\((base_libs|specific_w2rap-contigger) OBJECT/ { f = 1; next }
f { line = $0; closing = (line ~ /\)/); gsub(/[ \t\)]/, "", line);
    if (line ~ /^src.*\.cc$/) print line; if (closing) f = 0 }
' "$REF/CMakeLists.txt" | sort -u
echo src/modules/w2rap-contigger.cc
