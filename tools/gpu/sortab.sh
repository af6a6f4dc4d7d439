#!/bin/bash
# A/B of the sort alone: build/sort_plain (working tree) against build/sort_plain_old (reference build), time + instruction count
build/sort_plain_old | grep -E "cells|check" | paste - - | awk '{print "old", $2, $5, $6, $16, $17, $18, $19}'
build/sort_plain | grep -E "cells|check" | paste - - | awk '{print "new", $2, $5, $6, $16, $17, $18, $19}'
for b in build/sort_plain_old build/sort_plain; do ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:rs_sort_test -s 24 -c 1 $b 2>&1 | grep -E "inst_executed|issue_active" | awk -v b=$b '{print b, $1, $3}'; done
