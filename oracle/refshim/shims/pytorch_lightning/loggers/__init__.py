class LightningLoggerBase: pass
