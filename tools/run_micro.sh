for pair in -1; do
  echo "== VB_PAIR=$pair"
  VB_PAIR=$pair VB_EPI=simple,r1s timeout 150 python tools/conv_micro.py 2>&1 | tail -70
done
