#!/bin/bash
# SASS instruction count per kernel of a shared object (static code size = 16 bytes per instruction)
cuobjdump -sass "${1:-imitation-learning-rl_b200/libilrl_b200.so}" | awk '/Function : /{name=$3} /^ +\/\*[0-9a-f]+\*\/ /{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -rn
