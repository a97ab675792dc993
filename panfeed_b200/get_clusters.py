"""`panfeed-get-clusters` (/root/reference/panfeed/get_clusters.py): python -m panfeed_b200.get_clusters"""
from .postgwas import get_clusters_main as main

if __name__ == "__main__":
    main()
