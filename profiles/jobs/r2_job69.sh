python profiles/jobs/rough_split2.py 2>&1 | grep -v Warn
