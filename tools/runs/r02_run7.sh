#!/bin/bash
for v in endoinline endocalls small endocalls_small minb4 small_minb4; do echo "== $v"; timeout 120 ./tools/kexp/pa_$v; done
