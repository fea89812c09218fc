#!/bin/bash
# Build neuron_poker_b200/build/libnpk_checked.so: the library with -DNPK_CHECKED (bounds checks on every shared-memory gather,
# deck slot and decoded index, deck-restored check after every work item; csrc/npk_device.cuh).  Select it with NPK_LIBRARY.
cd "$(dirname "$0")/.." && python neuron_poker_b200/_build.py -DNPK_CHECKED -o$PWD/neuron_poker_b200/build/libnpk_checked.so
