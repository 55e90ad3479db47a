timeout 300 python profiles/panel_rows_probe.py 2>&1 | tail -12
echo ALLDONE_MARK22
