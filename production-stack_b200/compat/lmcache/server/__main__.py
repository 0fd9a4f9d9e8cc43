import sys

from b200kv.server import main

if __name__ == "__main__":
    sys.exit(main())
