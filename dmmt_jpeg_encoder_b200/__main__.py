import sys

from .reference_api import main

sys.exit(main())
