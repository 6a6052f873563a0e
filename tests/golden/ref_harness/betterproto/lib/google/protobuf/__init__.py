from datetime import datetime as Timestamp
