timeout 120 python tools/att_timeline.py 2>&1 | tail -36
ADA_ATT_PAD=1 timeout 120 python tools/att_timeline.py 2>&1 | tail -36
