"""`panfeed-get-kmers` (/root/reference/panfeed/get_kmers.py): python -m panfeed_b200.get_kmers"""
from .postgwas import get_kmers_main as main

if __name__ == "__main__":
    main()
