#!/bin/bash
# Per-opcode counts of the tensor-core / TMA / TMEM instructions in the built library (cuobjdump -sass), per kernel.
SO=${1:-pcgan_b200/libpcgan_kernels.so}
OUT=${2:-profiles/r2_sass_ops.txt}
cuobjdump -sass "$SO" > /tmp/pcgan_sass.txt
{
  echo "# cuobjdump -sass $SO | grep -c <opcode>   ($(date -u +%F), nvcc $(nvcc --version | grep -o "V[0-9][0-9.]*" | tail -1))"
  for op in UTCHMMA UTCHMMA.2CTA UTCQMMA LDTM STTM UTMALDG UTMALDG.5D UTMASTG UBLKCP UTCBAR UTCATOMSWS UTMAPF SYNCS.ARRIVE SYNCS.PHASECHK REDG RED.E ATOMG HMMA; do
    printf "%-16s %s\n" "$op" "$(grep -c "[^A-Z.]$op[ .]" /tmp/pcgan_sass.txt)"
  done
  echo "# per kernel: Function name, then counts of UTCHMMA / LDTM / UTMALDG / UTMASTG / UBLKCP"
  awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /UTMALDG/ {c[name]++} /UTMASTG/ {d[name]++} /UBLKCP/ {e[name]++} /Function :/ {n[name]=1}
       END {for (k in n) if (a[k]+b[k]+c[k]+d[k]+e[k] > 0) printf "%-90s UTCHMMA=%d LDTM=%d UTMALDG=%d UTMASTG=%d UBLKCP=%d\n", substr(k,1,90), a[k], b[k], c[k], d[k], e[k]}' /tmp/pcgan_sass.txt | sort
} > "$OUT"
cat "$OUT"
