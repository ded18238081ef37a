#!/bin/bash
# Build A/B variants of the library HERE (no GPU needed) into sings_b200/lib/variants/<name>.so;
# on the GPU box tools/ab_variants.sh benches each one via SGS_LIB_PATH.  One line per variant in $1:
#     name | make arguments (EXTRA="-D..." ...)
set -e
mkdir -p sings_b200/lib/variants
while IFS='|' read -r name margs; do
  name=$(echo $name | tr -d " +")
  [ -z "$name" ] && continue
  case "$name" in \#*) continue;; esac
  echo "=== $name :: $margs"
  rm -rf /tmp/sgs_variant_build && mkdir -p /tmp/sgs_variant_build
  eval make -s -j8 -C sings_b200/csrc OBJDIR=/tmp/sgs_variant_build OUT=../lib/variants/$name.so $margs > /tmp/make_variant.log 2>&1 || (tail -5 /tmp/make_variant.log; exit 1)
done < "$1"
ls -la sings_b200/lib/variants
