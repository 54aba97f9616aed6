tools/ab_variants.sh r2an "base|-|" "top9|top9|" "top73|top73|" "top73g|top73g|" "top200|top200|" "base2|-|"
