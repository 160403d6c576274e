L=$PWD/vit_deep_radiomics_b200/lib
one() { python bench.py --no-sub 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['attention']['achieved']))"; }
for v in p0x22 p0x2A p0x55; do VDR_LIB=$L/libvdr_$v.so python tools/attn_ab.py 1024 1025 2>&1 | tail -2; done
python tools/attn_ab.py 1024 1025 2>&1 | tail -2
for i in 1 2; do
  one base
  VDR_LIB=$L/libvdr_p0x2A.so one p3
  VDR_LIB=$L/libvdr_p0x22.so one p2
  VDR_LIB=$L/libvdr_p0x55.so one p4
done
