#!/bin/bash
# A/B of quint-kernel versions: decoder forward / latent-gradient time per call with each variant library
for l in ${LIBS:-head work b93bce5 head work}; do
  echo "== $l"
  OP3_B200_LIB=$PWD/op3_b200/_build/var/lib_$l.so timeout 120 python tools/dec_time.py 2>&1 | tail -2
done
