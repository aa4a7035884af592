"""Stand-in for matplotlib (not installed): /root/reference/dataloader/utils.py:3 imports pyplot at module level and
never uses it on the paths oracle/make_golden.py executes."""
