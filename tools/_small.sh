echo "== slab"; python tools/_small.py 2>&1 | tail -6
echo "== no slab"; W2RAP_NO_SLAB=1 python tools/_small.py 2>&1 | tail -6
echo "== slab traced"; W2RAP_TRACE=1 python tools/_small.py 2>&1 | tail -45
