from .controller import main
import sys
sys.exit(main())
