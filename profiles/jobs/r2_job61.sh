set -x
python profiles/jobs/stagger_sweep.py 32768 1100,555,296 1200,555,296 1100,592,296 1200,518,330 1300,518,296 1200,518,296
CASE=go1 python profiles/jobs/stagger_sweep.py 32768 default 1100,518,296 1200,518,296 1100,555,296
python profiles/jobs/stagger_sweep.py 24576 default 1100,518,296 1100,444,296
python profiles/jobs/stagger_sweep.py 20000 default 1100,518,296
