class ServiceBase: pass
