cd /root/repo
for so in build_variants/v*.so; do echo "== $so $(grep "^$(basename $so .so):" build_variants/list.txt)"; RMB200_LIB=$PWD/$so python tools/run_once.py --config 4 --users 37888 --reps 3 2>&1 | tail -1 | cut -c1-100; done
