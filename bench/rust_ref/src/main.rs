//! Reference replay of the committed fixtures (see Cargo.toml).  Every block cites the reference lines it restates;
//! nothing here is shared with oracle/ or pingoo_b200/.
//!
//! Fixture format (tests/golden/make_scenario_goldens.py): a JSON array of cases
//!   { name, eval_gates, rules: [{name, expression|null, actions: [1=block|2=captcha]}], services: [{name, route|null}],
//!     lists: {name: [type 0=String|1=Int|2=Ip, csv text]},
//!     requests: [{host,url,path,method,user_agent, ip_hex (16 bytes, IPv4 in the first 4), ip_is_v6, remote_port, asn,
//!                 country (two ASCII bytes, first letter in the low byte), flags}],
//!     verdicts: [action | rule << 2], services_out: [index | 0xFFFF] }
//! flags: 1 captcha_verified, 2 pre-blocked, 4 pre-captcha, 8 bypass (include/pingoo_waf.h).
use std::collections::HashMap;
use std::net::{IpAddr, Ipv4Addr, Ipv6Addr};
use std::str::FromStr;
use std::time::Instant;

use ipnetwork::IpNetwork;
use serde::{Deserialize, Serialize};

const NO_RULE: u32 = 0x3FFF_FFFF;
const NO_SERVICE: u32 = 0xFFFF;

#[derive(Deserialize, Serialize, Clone)]
struct FxRule { name: String, expression: Option<String>, actions: Vec<u8> }
#[derive(Deserialize, Serialize, Clone)]
struct FxService { name: String, route: Option<String> }
#[derive(Deserialize, Serialize, Clone)]
struct FxRequest {
    host: String, url: String, path: String, method: String, user_agent: String,
    ip_hex: String, ip_is_v6: u8, remote_port: i32, asn: i64, country: u32, flags: u32,
}
#[derive(Deserialize, Serialize, Clone)]
struct FxCase {
    name: String, eval_gates: bool, rules: Vec<FxRule>, services: Vec<FxService>,
    lists: HashMap<String, (u8, String)>, requests: Vec<FxRequest>, verdicts: Vec<u32>, services_out: Vec<u32>,
}

// pingoo/rules.rs:16-34 (the url is serialised with Display, the method with as_str: both plain strings here)
#[derive(Serialize)]
struct RequestData<'a> { host: &'a str, url: &'a str, path: &'a str, method: &'a str, user_agent: &'a str }
#[derive(Serialize)]
struct ClientData { ip: IpAddr, remote_port: i32, asn: i64, country: String }

struct Rule { expression: Option<bel::Program>, actions: Vec<u8> }

// pingoo/lists.rs:62-113 (csv: no headers, flexible, 1..=2 columns, column 0 trimmed) and :115-125
fn load_lists(lists: &HashMap<String, (u8, String)>) -> bel::Value {
    let mut out: HashMap<String, bel::Value> = HashMap::new();
    for (name, (ty, text)) in lists {
        let mut rd = csv::ReaderBuilder::new().has_headers(false).flexible(true).from_reader(text.as_bytes());
        let mut items: Vec<String> = Vec::new();
        for rec in rd.records() {
            let rec = rec.expect("csv");
            assert!(rec.len() >= 1 && rec.len() <= 2, "invalid number of columns");
            items.push(rec[0].trim().to_string());
        }
        let v: bel::Value = match ty {
            0 => items.into(),
            1 => items.iter().map(|s| s.parse::<i64>().expect("int")).collect::<Vec<i64>>().into(),
            _ => items.iter().map(|s| s.parse::<IpNetwork>().expect("ip network")).collect::<Vec<IpNetwork>>().into(),
        };
        out.insert(name.clone(), v);
    }
    out.into()
}

fn ip_of(r: &FxRequest) -> IpAddr {
    let b: Vec<u8> = (0..16).map(|i| u8::from_str_radix(&r.ip_hex[2 * i..2 * i + 2], 16).unwrap()).collect();
    if r.ip_is_v6 != 0 {
        let mut a = [0u8; 16];
        a.copy_from_slice(&b);
        IpAddr::V6(Ipv6Addr::from(a))
    } else {
        IpAddr::V4(Ipv4Addr::new(b[0], b[1], b[2], b[3]))
    }
}

// rules::compile_expression (rules/rules.rs:45-53): a rule whose expression does not compile refuses to start the proxy
// (pingoo/config/config.rs:255-269); the fixtures only hold compiling rules.
fn compile(e: &Option<String>) -> Option<bel::Program> {
    e.as_ref().map(|s| bel::Program::compile(s).unwrap_or_else(|err| panic!("Expression is not valid: {err}: {s}")))
}

// Rule::match_request (pingoo/rules.rs:36-52)
fn matches(p: &Option<bel::Program>, ctx: &bel::Context) -> bool {
    match p {
        None => true,
        Some(prog) => match prog.execute(ctx) { Ok(v) => v == true.into(), Err(_) => false },
    }
}

fn evaluate(case: &FxCase, rules: &[Rule], services: &[Option<bel::Program>], lists: &bel::Value, r: &FxRequest) -> (u32, u32) {
    // pre-rule gates (http_listener.rs:196-204, 222-236); flags carry what the listener decided from cookies
    if r.flags & 2 != 0 { return (1 | NO_RULE << 2, NO_SERVICE); }
    if case.eval_gates {
        if r.user_agent.is_empty() || r.user_agent.len() >= 256 { return (1 | NO_RULE << 2, NO_SERVICE); }
        if r.flags & 8 != 0 || r.path.starts_with("/__pingoo/captcha") { return (3 | NO_RULE << 2, NO_SERVICE); }
    } else if r.flags & 8 != 0 { return (3 | NO_RULE << 2, NO_SERVICE); }
    if r.flags & 4 != 0 { return (2 | NO_RULE << 2, NO_SERVICE); }
    let captcha_verified = r.flags & 1 != 0;
    let c = r.country;
    let client = ClientData {
        ip: ip_of(r), remote_port: r.remote_port, asn: r.asn,
        country: String::from_utf8_lossy(&[(c & 0xFF) as u8, ((c >> 8) & 0xFF) as u8]).to_string(),
    };
    let req = RequestData { host: &r.host, url: &r.url, path: &r.path, method: &r.method, user_agent: &r.user_agent };
    // http_listener.rs:239-249
    let mut ctx = bel::Context::default();
    let _ = ctx.add_variable("http_request", req);
    let _ = ctx.add_variable("client", &client);
    ctx.add_variable_from_value("lists", lists);
    // http_listener.rs:251-264
    for (i, rule) in rules.iter().enumerate() {
        if matches(&rule.expression, &ctx) {
            for a in &rule.actions {
                match a {
                    1 => return (1 | (i as u32) << 2, NO_SERVICE),
                    2 => if !captcha_verified { return (2 | (i as u32) << 2, NO_SERVICE) },
                    _ => {}
                }
            }
        }
    }
    // http_listener.rs:266-272, services/http_proxy_service.rs:84-95
    for (i, route) in services.iter().enumerate() {
        if matches(route, &ctx) { return (NO_RULE << 2, i as u32); }
    }
    (NO_RULE << 2, NO_SERVICE)
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() < 3 { eprintln!("usage: pingoo_rust_ref check|write|bench <fixture.json> [repeat]"); std::process::exit(2); }
    let text = std::fs::read_to_string(&args[2]).expect("fixture");
    let mut cases: Vec<FxCase> = serde_json::from_str(&text).expect("fixture json");
    let mut diffs = 0usize;
    let mut total = 0usize;
    let repeat: usize = args.get(3).and_then(|s| usize::from_str(s).ok()).unwrap_or(1);
    let t0 = Instant::now();
    for case in cases.iter_mut() {
        let rules: Vec<Rule> = case.rules.iter().map(|r| Rule { expression: compile(&r.expression), actions: r.actions.clone() }).collect();
        let services: Vec<Option<bel::Program>> = case.services.iter().map(|s| compile(&s.route)).collect();
        let lists = load_lists(&case.lists);
        for _ in 0..repeat {
            for (i, r) in case.requests.clone().iter().enumerate() {
                let (v, s) = evaluate(case, &rules, &services, &lists, r);
                total += 1;
                let s_expect = if case.services.is_empty() { NO_SERVICE } else { case.services_out[i] };
                let s_got = if case.services.is_empty() { NO_SERVICE } else { s };
                if args[1] == "write" { case.verdicts[i] = v; if !case.services.is_empty() { case.services_out[i] = s; } }
                else if v != case.verdicts[i] || s_got != s_expect {
                    diffs += 1;
                    if args[1] == "check" {
                        println!("{} request {}: reference verdict {:#x} service {:#x}, fixture {:#x} / {:#x}", case.name, i, v, s_got, case.verdicts[i], s_expect);
                    }
                }
            }
        }
    }
    let dt = t0.elapsed().as_secs_f64();
    match args[1].as_str() {
        "write" => { std::fs::write(&args[2], serde_json::to_string(&cases).unwrap()).unwrap(); println!("rewrote {} verdicts", total); }
        "bench" => println!("{} evaluations in {:.3} s = {:.1} k req/s on one thread ({} differences)", total, dt, total as f64 / dt / 1e3, diffs),
        _ => { println!("{} evaluations, {} differences", total, diffs); if diffs > 0 { std::process::exit(1); } }
    }
}
